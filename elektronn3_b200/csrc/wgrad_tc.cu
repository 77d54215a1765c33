// Convolution weight gradient on tcgen05 tensor cores, sm_100a.
//
//   dW[tap][ci][co] = sum over output voxels v of  x[v + tap - pad][ci] * dy[v][co]
//
// (the backward-filter of nn.Conv3d at elektronn3 models/unet.py:131-149; with one tap on
// (x, space-to-depth(dy)) also nn.ConvTranspose3d's, unet.py:152-165).
//
// GEMM view: M = input channels, N = output channels, K = voxels.  Both operands are read K-major from
// Z-PLANAR fp16 copies of the tensors (N, D, C, H, Wp), Wp = ceil8(W) (kind::f16, fp32 accumulate; fp16
// carries TF32's 10 mantissa bits, the gradient copies carry the power-of-two scale of e3b_norm_bwd_*, undone
// in the split-K reduction): a TMA box of 64 consecutive x-voxels x CB channels x rows lands in shared
// memory as the canonical 128-byte-swizzled K-major tile (one 128 B line = 64 voxels of one channel;
// 8 channels = one 1024 B swizzle atom).
//  * the (kw) x-shifts of the stencil cannot be TMA coordinates (a box must start 16-byte aligned: a
//    1-voxel shift of a 2-byte element is an illegal instruction) nor descriptor offsets, so the kernel
//    that produces dy also writes it as kw x-shifted copies (N, D, kw, Co, H, Wp); the copies are stacked
//    in the MMA's N dimension: one MMA of N = kw*NTW columns serves all x taps;
//  * the (kh) y-shifts are STACKED IN M: a stage holds the rows [y][c][32 vox] contiguously, so an
//    M = 128 operand starting at row r covers rows r .. r+RS-1 (RS = 128 / CB, CB = 8/16/32 channels per
//    unit): accumulator rows [j*CB, (j+1)*CB) are tap dy = j.  With CB = 32 one MMA serves all three dy
//    taps (75 % of the M rows useful instead of 25 %);
//  * the (kd) z-shifts pair the x plane z with the dy planes z+pd-kd+1 .. z+pd: either all inside one CTA
//    (narrow N) or split over CTAs.  A CTA walks a contiguous run of x planes of one (n, y tile, x tile)
//    column, so the dy planes live in their own shared-memory ring and every plane is staged ONCE for
//    the kd x-planes that use it (a sliding window along z) instead of once per x plane;
//  * every (kd, dx) tap pair owns a TMEM accumulator (columns <= 512); the contraction over voxels is
//    split over CTAs (split-K), partials go to a workspace and a deterministic second kernel reduces
//    them into the torch weight layout.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace e3b {

static constexpr int kWgThreads = 192;
static constexpr int kSeg = 64;              // voxels per 128-byte line

struct WgradParams {
    int N, D, H, W;              // x extents (source 0 / the cropped view of source 1)
    int D1;                      // planes per sample in the allocation of source 1
    int Do, Ho, Wo;              // dy extents
    int kd, kh, kw, pd, ph, pw;
    int off1_d, off1_h, off1_w;
    int CB, RS;                  // channels per M unit, rows stacked in M
    int TY, TYA, rows_alloc;     // dy rows per stage, x rows loaded, x rows addressed by the MMAs
    int tiles_x, tiles_y;
    int mchunks0, mchunks;       // CB-channel chunks in source 0 / total
    int NTW, nchunks_n;          // N columns per CTA, number of N chunks
    int kdn, kd_units;           // kd taps per CTA, CTAs along kd
    int units, S, SA, SB;        // CTAs = units * S; depth of the x-tile ring and of the dy-plane ring
    int total_L;                 // linear work items: (column = (n, y tile, x tile)) x (x plane z)
    uint32_t a_load_bytes, a_bytes, b_plane_bytes;
    int ktot, npad_total;        // partial-sum row / column space
    float* part;                 // [S][ntaps][ktot][npad_total]
};

struct WgUnit { int ku, mc, nc; };

E3B_DEVINL WgUnit decode_unit(const WgradParams& p, int u) {
    WgUnit r;
    r.nc = u % p.nchunks_n; u /= p.nchunks_n;
    r.mc = u % p.mchunks; u /= p.mchunks;
    r.ku = u;
    return r;
}

// column index -> sample and tile origin
E3B_DEVINL void decode_col(const WgradParams& p, int col, int& n, int& y0, int& x0) {
    int xt = col % p.tiles_x; col /= p.tiles_x;
    int yt = col % p.tiles_y;
    n = col / p.tiles_y;
    x0 = xt * kSeg; y0 = yt * p.TY;
}

// The run of work of one CTA, cut into per-column segments.  Producer and MMA issuer walk it in lock step.
struct WgSeg {
    int col, za, zb;             // x planes [za, zb) of column col
    int blo, bhi;                // dy planes [blo, bhi] staged for this segment (may be empty: blo > bhi)
};
// dy planes used by x plane zi: [lo, hi]  (all kdn taps of this CTA)
E3B_DEVINL int wg_lo(const WgradParams& p, int ku, int zi) { return zi + p.pd - (ku * p.kdn + p.kdn - 1); }
E3B_DEVINL int wg_hi(const WgradParams& p, int ku, int zi) { return zi + p.pd - ku * p.kdn; }
E3B_DEVINL WgSeg wg_segment(const WgradParams& p, int ku, int L, int L1) {
    WgSeg g;
    g.col = L / p.D; g.za = L - g.col * p.D;
    const int left = L1 - L;
    g.zb = g.za + left < p.D ? g.za + left : p.D;
    g.blo = wg_lo(p, ku, g.za); if (g.blo < 0) g.blo = 0;
    g.bhi = wg_hi(p, ku, g.zb - 1); if (g.bhi > p.Do - 1) g.bhi = p.Do - 1;
    return g;
}

// K-major, 128-byte swizzle: 8-row groups 1024 B apart; the K advance inside a line is added to the start
// address (tiles are 1024 B aligned, so the swizzle phase bits [7,10) are untouched by offsets < 128 B).
E3B_DEVINL uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                       // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // SBO
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmx0, const __grid_constant__ CUtensorMap tmx1,
                const __grid_constant__ CUtensorMap tmdy, const WgradParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128 B swizzle atoms
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + (size_t)p.SA * p.a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)p.SB * p.b_plane_bytes);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + p.SA;
    uint64_t* b_full = a_empty + p.SA;
    uint64_t* b_empty = b_full + p.SB;
    uint64_t* done = b_empty + p.SB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    volatile uint32_t* started_slot = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x / p.S, split = blockIdx.x % p.S;
    const WgUnit u = decode_unit(p, unit);
    const bool src1 = u.mc >= p.mchunks0;
    const int mc_local = src1 ? u.mc - p.mchunks0 : u.mc;
    // this CTA's contiguous run of the linear (column, z) work space
    const int L0 = (int)(((long long)p.total_L * split) / p.S), L1 = (int)(((long long)p.total_L * (split + 1)) / p.S);

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.SA; i++) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < p.SB; i++) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(done, 2);                      // tcgen05.commit + the issuer's own (releasing) arrive
        fence_barrier_init();
        tma_prefetch_desc(src1 ? &tmx1 : &tmx0);
        tma_prefetch_desc(&tmdy);
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
            const CUtensorMap* mx = src1 ? &tmx1 : &tmx0;
            const int ox = src1 ? p.off1_w : 0, oy = src1 ? p.off1_h : 0, oz = src1 ? p.off1_d : 0;
            const int Dsrc = src1 ? p.D1 : p.D;      // planes per sample in the source's allocation
            for (int L = L0; L < L1;) {
                const WgSeg g = wg_segment(p, u.ku, L, L1);
                int n, y0, x0;
                decode_col(p, g.col, n, y0, x0);
                int bnext = g.blo;
                for (int zi = g.za; zi < g.zb; zi++) {
                    int need = wg_hi(p, u.ku, zi); if (need > g.bhi) need = g.bhi;
                    for (; bnext <= need; bnext++) {
                        // dy: dims (x, co, dxi, y, n*Do + z)
                        mbar_wait(&b_empty[sb], pb ^ 1);
                        mbar_arrive_expect_tx(&b_full[sb], p.b_plane_bytes);
                        tma_load_5d(b_ring + (size_t)sb * p.b_plane_bytes, &tmdy, &b_full[sb], x0, u.nc * p.NTW, 0, y0,
                                    n * p.Do + bnext);
                        if (++sb == (uint32_t)p.SB) { sb = 0; pb ^= 1; }
                    }
                    // x: dims (x, c, y, n*D + z)
                    mbar_wait(&a_empty[sa], pa ^ 1);
                    mbar_arrive_expect_tx(&a_full[sa], p.a_load_bytes);
                    tma_load_4d(a_ring + (size_t)sa * p.a_bytes, mx, &a_full[sa], x0 + ox, mc_local * p.CB, y0 - p.ph + oy,
                                n * Dsrc + zi + oz);
                    if (++sa == (uint32_t)p.SA) { sa = 0; pa ^= 1; }
                }
                L += g.zb - g.za;
            }
        }
    } else if (warp == 1) {
        // whole warp runs the loop (uniform control flow); one elected lane issues
        const uint32_t idesc = umma_idesc_f16(p.kw * p.NTW, 0, 0);
        const bool leader = elect_one();
        uint32_t sa = 0, pa = 0;
        uint32_t wb = 0, wpb = 0;                    // next dy-plane slot to wait for
        uint32_t rb = 0;                             // next dy-plane slot to release
        uint32_t started = 0;                        // bit kj: accumulator kj has been written once
        const uint32_t a_row16 = (uint32_t)p.CB * 8u, b_row16 = (uint32_t)(p.kw * p.NTW) * 8u;   // 16-byte units
        const uint64_t tmpl = umma_desc_sw128(0);
        const uint32_t a16 = smem_u32(a_ring) >> 4, b16 = smem_u32(b_ring) >> 4;
        for (int L = L0; L < L1;) {
            const WgSeg g = wg_segment(p, u.ku, L, L1);
            const uint32_t slot0 = wb;               // ring slot of dy plane g.blo
            int bwaited = g.blo, brel = g.blo;
            for (int zi = g.za; zi < g.zb; zi++) {
                int need = wg_hi(p, u.ku, zi); if (need > g.bhi) need = g.bhi;
                for (; bwaited <= need; bwaited++) {
                    mbar_wait(&b_full[wb], wpb);
                    if (++wb == (uint32_t)p.SB) { wb = 0; wpb ^= 1; }
                }
                mbar_wait(&a_full[sa], pa);
                tc_fence_after();
                const uint32_t sA16 = a16 + (uint32_t)sa * (p.a_bytes >> 4);
                for (int kj = 0; kj < p.kdn; kj++) {
                    const int z = zi + p.pd - (u.ku * p.kdn + kj);
                    if (z < 0 || z >= p.Do) continue;
                    if (leader) {
                        const uint32_t slot = (slot0 + (uint32_t)(z - g.blo)) % (uint32_t)p.SB;
                        const uint32_t acc = tmem_base + (uint32_t)(kj * p.kw * p.NTW);
                        uint64_t ad = tmpl + sA16;
                        uint64_t bd = tmpl + (b16 + slot * (p.b_plane_bytes >> 4));
                        const uint32_t first = ((started >> kj) & 1u) ? 1u : 0u;
                        for (int yy = 0; yy < p.TY; yy++) {
                            umma_f16(acc, ad, bd, idesc, first | (uint32_t)yy);
                            umma_f16(acc, ad + 2, bd + 2, idesc, 1u);
                            umma_f16(acc, ad + 4, bd + 4, idesc, 1u);
                            umma_f16(acc, ad + 6, bd + 6, idesc, 1u);
                            ad += a_row16; bd += b_row16;
                        }
                    }
                    started |= 1u << kj;
                }
                if (leader) umma_commit(&a_empty[sa]);
                if (++sa == (uint32_t)p.SA) { sa = 0; pa ^= 1; }
                // dy planes no later x plane of this segment uses are handed back to the producer
                int relupto = g.bhi + 1;
                if (zi + 1 < g.zb) { relupto = wg_lo(p, u.ku, zi + 1); if (relupto > g.bhi + 1) relupto = g.bhi + 1; }
                for (; brel < relupto; brel++) {
                    if (leader) umma_commit(&b_empty[rb]);
                    if (++rb == (uint32_t)p.SB) rb = 0;
                }
                __syncwarp();
            }
            L += g.zb - g.za;
        }
        if (leader) {
            // accumulators that never received a plane (tiny volumes) are reported to the epilogue as empty
            *started_slot = started;
            mbar_arrive(done);
            umma_commit(done);
        }
        __syncwarp();
    } else {
        // epilogue: TMEM -> split-K partial.  Accumulator row r = j*CB + c  <->  tap dy = j, channel c.
        mbar_wait(done, 0);
        tc_fence_after();
        const uint32_t started = *started_slot;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int j = row / p.CB, c = row % p.CB;
        const int ntaps = p.kd * p.kh * p.kw;
        const bool row_ok = j < p.kh;
        for (int kj = 0; kj < p.kdn; kj++) {
            const int kdi = u.ku * p.kdn + kj;
            const bool any = (started >> kj) & 1u;
            for (int dxi = 0; dxi < p.kw; dxi++) {
                const int tap = (kdi * p.kh + j) * p.kw + dxi;
                for (int cb = 0; cb < p.NTW; cb += 16) {
                    float v[16];
                    if (any) {
                        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((kj * p.kw + dxi) * p.NTW + cb), v);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; i++) v[i] = 0.f;
                    }
                    if (row_ok) {
                        float* o = p.part + (((size_t)split * ntaps + tap) * p.ktot + (size_t)u.mc * p.CB + c) * p.npad_total +
                                   u.nc * p.NTW + cb;
#pragma unroll
                        for (int j4 = 0; j4 < 4; j4++)
                            reinterpret_cast<float4*>(o)[j4] = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// deterministic split-K reduction + scatter into the torch parameter layout.  A block of 8 warps owns 32 consecutive
// elements: warp w adds the partials s = w, w + 8, ... (coalesced 128-byte rows), the eight sums meet in shared
// memory and are added in warp order.  (One thread per element walking all S <= 148 partials was latency bound:
// 25 us for the 16 MB of partials of a 32 -> 32 layer.)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int S, int ntaps,
                                                            int ktot, int npad_total, int CB, int mchunks0, int C0, int C1, int Co,
                                                            int layout, int up_taps, int up_co, int up_copad,
                                                            const float* __restrict__ dy_unscale)
{
    __shared__ float red[8][32];
    const size_t total = (size_t)ntaps * ktot * npad_total;
    const float unscale = dy_unscale ? __ldg(dy_unscale) : 1.f;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (size_t i0 = (size_t)blockIdx.x * 32; i0 < total; i0 += (size_t)gridDim.x * 32) {
        const size_t i = i0 + lane;
        bool valid = i < total;
        int nn = 0, ci = 0, tap = 0;
        if (valid) {
            nn = (int)(i % npad_total);
            const int kk = (int)((i / npad_total) % ktot);
            tap = (int)(i / ((size_t)npad_total * ktot));
            if (kk < mchunks0 * CB) { valid = kk < C0; ci = kk; }
            else { const int k1 = kk - mchunks0 * CB; valid = k1 < C1; ci = C0 + k1; }
            if (layout == 0) valid = valid && nn < Co;
            else valid = valid && (nn / up_copad < up_taps) && (nn % up_copad < up_co);
        }
        float s = 0.f;
        if (valid) for (int sp = w; sp < S; sp += 8) s += part[(size_t)sp * total + i];
        red[w][lane] = s;
        __syncthreads();
        if (w == 0 && valid) {
            float t = red[0][lane];
#pragma unroll
            for (int j = 1; j < 8; j++) t += red[j][lane];
            t *= unscale;
            if (layout == 0) dw[((size_t)nn * (C0 + C1) + ci) * ntaps + tap] = t;
            else dw[((size_t)ci * up_co + nn % up_copad) * up_taps + nn / up_copad] = t;
        }
        __syncthreads();
    }
}

// one thread per element, walking all S partials: for layers whose contraction is split over fewer than 64 CTAs
__global__ void wgrad_reduce_flat_kernel(const float* __restrict__ part, float* __restrict__ dw, int S, int ntaps, int ktot,
                                         int npad_total, int CB, int mchunks0, int C0, int C1, int Co, int layout, int up_taps,
                                         int up_co, int up_copad, const float* __restrict__ dy_unscale)
{
    const size_t total = (size_t)ntaps * ktot * npad_total;
    const float unscale = dy_unscale ? __ldg(dy_unscale) : 1.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int nn = (int)(i % npad_total);
        const int kk = (int)((i / npad_total) % ktot);
        const int tap = (int)(i / ((size_t)npad_total * ktot));
        int ci;
        if (kk < mchunks0 * CB) { if (kk >= C0) continue; ci = kk; }
        else { const int k1 = kk - mchunks0 * CB; if (k1 >= C1) continue; ci = C0 + k1; }
        if (layout == 0) { if (nn >= Co) continue; }
        else { if (nn / up_copad >= up_taps || nn % up_copad >= up_co) continue; }
        float s = 0.f;
        for (int sp = 0; sp < S; sp++) s += part[(size_t)sp * total + i];
        s *= unscale;
        if (layout == 0) dw[((size_t)nn * (C0 + C1) + ci) * ntaps + tap] = s;
        else dw[((size_t)ci * up_co + nn % up_copad) * up_taps + nn / up_copad] = s;
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_enc()
{
    static PFN_encodeTiled enc = nullptr;
    if (!enc) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        enc = reinterpret_cast<PFN_encodeTiled>(fp);
    }
    return enc;
}

// z-planar fp16 activation (N, D, C, H, Wp) viewed as 4D (x: W, c: C, y: H, zn: N*D); box (64, bc, by, 1), 128 B swizzle
static int make_x_map(CUtensorMap* map, const void* ptr, int N, int C, int D, int H, int W, int bc, int by)
{
    PFN_encodeTiled enc = get_enc();
    if (!enc) return set_error("cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t Wp = (cuuint64_t)((W + 7) & ~7);
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)N * D};
    cuuint64_t strides[3] = {(cuuint64_t)H * Wp * 2, Wp * 2, (cuuint64_t)C * H * Wp * 2};
    cuuint32_t box[4] = {(cuuint32_t)kSeg, (cuuint32_t)bc, (cuuint32_t)by, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (box[1] > 256 || box[2] > 256) return set_error("wgrad: TMA box dimension > 256");
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("wgrad: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
    return 0;
}

// shifted gradient copies (N, Do, kw, Co, Ho, Wxp) viewed as 5D (x: Wx, co: Co, dxi: kw, y: Ho, zn: N*Do);
// box (32, NTW, kw, TY, 1)
static int make_dy_map(CUtensorMap* map, const void* ptr, int N, int Co, int kw, int Do, int Ho, int Wx, int bn, int by)
{
    PFN_encodeTiled enc = get_enc();
    if (!enc) return set_error("cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t Wp = (cuuint64_t)((Wx + 7) & ~7);
    cuuint64_t dims[5] = {(cuuint64_t)Wx, (cuuint64_t)Co, (cuuint64_t)kw, (cuuint64_t)Ho, (cuuint64_t)N * Do};
    cuuint64_t strides[4] = {(cuuint64_t)Ho * Wp * 2, (cuuint64_t)Co * Ho * Wp * 2, Wp * 2, (cuuint64_t)kw * Co * Ho * Wp * 2};
    cuuint32_t box[5] = {(cuuint32_t)kSeg, (cuuint32_t)bn, (cuuint32_t)kw, (cuuint32_t)by, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (box[1] > 256 || box[3] > 256) return set_error("wgrad: TMA box dimension > 256");
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("wgrad: cuTensorMapEncodeTiled(dy) failed (%d)", (int)r);
    return 0;
}

static int plan_wgrad(const e3b_wgrad_args* a, WgradParams& p)
{
    memset(&p, 0, sizeof(p));
    p.N = a->N; p.D = a->D; p.H = a->H; p.W = a->W;
    p.D1 = a->D1;
    p.kd = a->kd; p.kh = a->kh; p.kw = a->kw; p.pd = a->pd; p.ph = a->ph; p.pw = a->pw;
    p.Do = a->D + 2 * a->pd - a->kd + 1; p.Ho = a->H + 2 * a->ph - a->kh + 1; p.Wo = a->W + 2 * a->pw - a->kw + 1;
    if (p.Do <= 0 || p.Ho <= 0 || p.Wo <= 0) return set_error("wgrad: empty output");
    p.off1_d = a->off1_d; p.off1_h = a->off1_h; p.off1_w = a->off1_w;
    const int C1 = a->src1 ? a->C1 : 0;
    const int cmax = a->C0 > C1 ? a->C0 : C1;
    p.CB = cmax > 16 ? 32 : (cmax > 8 ? 16 : 8);
    p.RS = 128 / p.CB;
    p.mchunks0 = (a->C0 + p.CB - 1) / p.CB;
    p.mchunks = p.mchunks0 + (C1 + p.CB - 1) / p.CB;
    p.ktot = p.mchunks * p.CB;
    // N columns per CTA: one MMA covers the kw shifted copies, N = kw*NTW <= 256
    const int npad = cpad16(a->Co);
    const int ntw_max = a->kw == 1 ? 128 : 80;
    p.NTW = npad <= ntw_max ? npad : (npad % 64 == 0 ? 64 : (npad % 48 == 0 ? 48 : 32));
    p.nchunks_n = (npad + p.NTW - 1) / p.NTW;
    p.npad_total = p.nchunks_n * p.NTW;
    p.kdn = (a->kd * a->kw * p.NTW <= 512) ? a->kd : 1;
    p.kd_units = a->kd / p.kdn;
    p.units = p.kd_units * p.mchunks * p.nchunks_n;
    const size_t budget = 227 * 1024 - 2048 - 512;
    int ty = 8;
    // developer overrides for tuning runs (scripts/layer_bench.py): E3B_WGRAD_TY / _SA / _SB
    const char* e_ty = getenv("E3B_WGRAD_TY"); const char* e_sa = getenv("E3B_WGRAD_SA"); const char* e_sb = getenv("E3B_WGRAD_SB");
    if (e_ty) ty = atoi(e_ty);
    for (;; ty >>= 1) {
        p.TY = ty; p.TYA = ty + a->kh - 1; p.rows_alloc = ty + p.RS - 1;
        if (p.rows_alloc < p.TYA) p.rows_alloc = p.TYA;
        p.a_load_bytes = (uint32_t)(p.TYA * p.CB * 128);
        p.a_bytes = (uint32_t)(p.rows_alloc * p.CB * 128);
        p.b_plane_bytes = (uint32_t)(ty * a->kw * p.NTW * 128);
        // rings: the dy window holds kdn live planes; +2 lets the producer run ahead.  x tiles: 2..4 deep.
        const bool small_enough = ty == 1 || ty / 2 < p.Ho;      // do not carry rows a small volume does not have
        int sa = 2, sb = p.kdn + 1;
        const bool fits = (size_t)sa * p.a_bytes + (size_t)sb * p.b_plane_bytes <= budget;
        if (fits && (small_enough || ty == 1)) {
            for (;;) {
                bool grew = false;
                if (sb < p.kdn + 3 && (size_t)sa * p.a_bytes + (size_t)(sb + 1) * p.b_plane_bytes <= budget) { sb++; grew = true; }
                if (sa < 4 && (size_t)(sa + 1) * p.a_bytes + (size_t)sb * p.b_plane_bytes <= budget) { sa++; grew = true; }
                if (!grew) break;
            }
            p.SA = sa; p.SB = sb;
            if (e_sa && atoi(e_sa) >= 2 && atoi(e_sa) <= sa) p.SA = atoi(e_sa);
            if (e_sb && atoi(e_sb) >= p.kdn + 1 && atoi(e_sb) <= sb) p.SB = atoi(e_sb);
            break;
        }
        if (ty == 1) return set_error("wgrad: stage does not fit shared memory");
    }
    // tiles cover the conv INPUT width (the shifted gradient copies are indexed by the input x)
    p.tiles_x = (a->W + kSeg - 1) / kSeg; p.tiles_y = (p.Ho + p.TY - 1) / p.TY;
    p.total_L = p.tiles_x * p.tiles_y * a->N * p.D;
    int S = num_sms() / p.units; if (S < 1) S = 1;             // one resident CTA per SM, a single wave
    if (S > p.total_L) S = p.total_L;
    p.S = S;
    return 0;
}

int64_t wgrad_workspace_floats(const e3b_wgrad_args* a)
{
    WgradParams p;
    if (plan_wgrad(a, p)) return -1;
    return (int64_t)p.S * a->kd * a->kh * a->kw * p.ktot * p.npad_total;
}

int launch_wgrad_tc(const e3b_wgrad_args* a, cudaStream_t stream)
{
    WgradParams p;
    int rc = plan_wgrad(a, p);
    if (rc) return rc;
    p.part = a->workspace;
    CUtensorMap mx0, mx1, mdy;
    rc = make_x_map(&mx0, a->src0, a->N, a->C0, a->D, a->H, a->W, p.CB, p.TYA);
    if (rc) return rc;
    if (a->src1) {
        if ((a->off1_d | a->off1_h | a->off1_w) && (a->pd | a->ph | a->pw))
            return set_error("wgrad: a centre-cropped second source requires zero padding (VALID convolution)");
        if (a->off1_w & 7) return set_error("wgrad: the x crop offset of the second source must be a multiple of 8");
        rc = make_x_map(&mx1, a->src1, a->N, a->C1, a->D1, a->H1, a->W1, p.CB, p.TYA);
        if (rc) return rc;
    } else mx1 = mx0;
    rc = make_dy_map(&mdy, a->dy, a->N, a->Co, a->kw, p.Do, p.Ho, a->W, p.NTW, p.TY);
    if (rc) return rc;
    const size_t smem = (size_t)p.SA * p.a_bytes + (size_t)p.SB * p.b_plane_bytes + 1024 + 1024;
    static bool configured[kMaxDevices] = {false};
    if (!configured[current_device()]) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured[current_device()] = true;
    }
    wgrad_tc_kernel<<<p.units * p.S, kWgThreads, smem, stream>>>(mx0, mx1, mdy, p);
    rc = check_launch("wgrad_tc");
    if (rc) return rc;
    const int ntaps = a->kd * a->kh * a->kw;
    const size_t total = (size_t)ntaps * p.ktot * p.npad_total;
    if (p.S >= 64) {
        int blocks = (int)((total + 31) / 32); if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
        wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(p.part, a->dw, p.S, ntaps, p.ktot, p.npad_total, p.CB, p.mchunks0,
                                                        a->C0, a->src1 ? a->C1 : 0, a->Co, a->layout, a->up_taps, a->up_co,
                                                        cpad8(a->up_co), a->dy_unscale);
    } else {
        int blocks = (int)((total + 255) / 256); if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
        wgrad_reduce_flat_kernel<<<blocks, 256, 0, stream>>>(p.part, a->dw, p.S, ntaps, p.ktot, p.npad_total, p.CB, p.mchunks0,
                                                             a->C0, a->src1 ? a->C1 : 0, a->Co, a->layout, a->up_taps, a->up_co,
                                                             cpad8(a->up_co), a->dy_unscale);
    }
    return check_launch("wgrad_reduce");
}

}  // namespace e3b
