// Convolution weight gradient on tcgen05 tensor cores, sm_100a.
//
//   dW[tap][ci][co] = sum over output voxels v of  x[v + tap - pad][ci] * dy[v][co]
//
// (the backward-filter of nn.Conv3d at elektronn3 models/unet.py:131-149; with one tap on
// (x, space-to-depth(dy)) also nn.ConvTranspose3d's, unet.py:152-165).
//
// GEMM view: M = input channels, N = output channels, K = voxels.  Both operands are read STRAIGHT from the QH operand
// tensors the forward / dgrad kernels use (fp16, 8 channels per 16-byte voxel unit) as MN-major no-swizzle UMMA operands:
// 8 consecutive x-voxels of an 8-channel plane are one canonical core matrix (8 K-rows of 16 bytes), so one kind::f16
// MMA (K = 16) contracts 16 consecutive x-voxels, and -- because a voxel is a 16-byte unit -- every stencil shift is a
// legal descriptor start address.  No K-major ("planar") copies of the activations or of the gradients exist any more:
// round 1 wrote one planar copy of every activation and three x-shifted copies of every gradient (a TMA box of a
// 2-byte type cannot start at an odd element, profiles/r02_tma_odd_coordinate.txt), 268 MB of extra HBM traffic per
// full-resolution 32-channel layer.
//  * x taps (kw): the x tile carries a halo, tap dx is a +16 B start address; one TMEM accumulator per dx;
//  * y taps (kh) are STACKED IN M: the x tile lies in shared memory as [row][8-channel plane][x], so the 16 M groups of
//    one MMA are (row j, plane): accumulator rows [j*CB, (j+1)*CB) are tap dy = j of the gradient row they meet
//    (CB = 32 channels: 3 of the 4 stacked rows are useful);
//  * z taps (kd) are STACKED IN N: the gradient tile holds the kd planes z+pd-kd+1 .. z+pd as [row][plane z][channel
//    plane][x] (one TMA box; planes outside the volume are zero-filled = the convolution's zero padding), so the N groups
//    of one MMA are (z plane, channel plane): N = kd * NTW columns, accumulator column block zs is tap dz = kd-1-zs.
// One MMA of M = 128, N = 96, K = 16 therefore serves 9 taps of a 32 -> 32 layer; the contraction over voxels is split
// over CTAs (split-K), partials go to a workspace and a deterministic second kernel reduces them into the torch layout.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace e3b {

static constexpr int kWgThreads = 192;

struct WgradParams {
    int N, D, H, W;              // x extents (source 0 / the cropped view of source 1)
    int Do, Ho, Wo;              // dy extents
    int kd, kh, kw, pd, ph, pw;
    int CB, RS, PA, PB;          // channels per M block, rows stacked in M, 8-channel planes per x tile / per dy z-slot
    int TY, TX, HXa, rows_a;     // dy rows and x-voxels per step, x tile width (TX + kw - 1) and rows (TY + RS - 1)
    int tiles_x, tiles_y;
    int mchunks0, mchunks;       // CB-channel chunks in source 0 / total
    int NTW, nchunks_n;          // N columns per z-slot and CTA, number of N chunks
    int units, S, SA;            // CTAs = units * S; stage ring depth
    int total_L;                 // linear work items: (column = (n, y tile, x tile)) x (x plane z)
    uint32_t a_bytes, b_bytes, stage_bytes, tx_bytes;   // a_bytes: padded offset of the dy tile; tx_bytes: what the two TMA boxes move
    int ktot, npad_total;        // partial-sum row / column space
    float* part;                 // [S][ntaps][ktot][npad_total]
};

struct WgUnit { int mc, nc; };

E3B_DEVINL WgUnit decode_unit(const WgradParams& p, int u) {
    WgUnit r;
    r.nc = u % p.nchunks_n;
    r.mc = u / p.nchunks_n;
    return r;
}

// linear work item -> sample, tile origin, x plane
E3B_DEVINL void decode_item(const WgradParams& p, int L, int& n, int& y0, int& x0, int& zi) {
    int col = L / p.D;
    zi = L - col * p.D;
    const int xt = col % p.tiles_x; col /= p.tiles_x;
    const int yt = col % p.tiles_y;
    n = col / p.tiles_y;
    x0 = xt * p.TX; y0 = yt * p.TY;
}

// kind::f16 MMA with the descriptors given as (low, high) words (32-bit uniform adds advance the low words)
E3B_DEVINL void umma_f16_lohi_acc(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                  uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// the MMAs of one dy row: KW x taps (one accumulator each) x KS 16-voxel K steps, fully unrolled so that the UTCHMMAs issue
// back to back (a loop with run-time bounds costs the single issuing thread more than an MMA takes to execute)
template <int KW, int KS>
E3B_DEVINL void wg_issue_row(uint32_t tmem_base, uint32_t acc_cols, uint32_t a_row, uint32_t a_hi, uint32_t b_row, uint32_t b_hi,
                             uint32_t idesc, uint32_t acc_flag)
{
#pragma unroll
    for (int dx = 0; dx < KW; dx++) {
#pragma unroll
        for (int k = 0; k < KS; k++)
            umma_f16_lohi_acc(tmem_base + (uint32_t)dx * acc_cols, a_row + (uint32_t)(dx + 16 * k), a_hi, b_row + (uint32_t)(16 * k), b_hi,
                              idesc, k == 0 ? acc_flag : 1u);
    }
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmx0, const __grid_constant__ CUtensorMap tmx1,
                const __grid_constant__ CUtensorMap tmdy, const WgradParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.SA * p.stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = full + p.SA;
    uint64_t* done = empty + p.SA;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x / p.S, split = blockIdx.x % p.S;
    const WgUnit u = decode_unit(p, unit);
    const bool src1 = u.mc >= p.mchunks0;
    const int mc_local = src1 ? u.mc - p.mchunks0 : u.mc;
    // this CTA's contiguous run of the linear (column, z) work space
    const int L0 = (int)(((long long)p.total_L * split) / p.S), L1 = (int)(((long long)p.total_L * (split + 1)) / p.S);

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.SA; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
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
        // ===================== TMA producer: one x tile + one dy tile (kd planes) per stage =====================
        if (lane == 0) {
            uint32_t s = 0, par = 0;
            const CUtensorMap* mx = src1 ? &tmx1 : &tmx0;
            for (int L = L0; L < L1; L++) {
                int n, y0, x0, zi;
                decode_item(p, L, n, y0, x0, zi);
                mbar_wait(&empty[s], par ^ 1);
                mbar_arrive_expect_tx(&full[s], p.tx_bytes);
                uint8_t* st = smem + (size_t)s * p.stage_bytes;
                // x : dims (x*4, channel plane, y, z, n)       box (HXa*4, PA, rows_a, 1, 1)
                tma_load_5d(st, mx, &full[s], (x0 - p.pw) * 4, mc_local * p.PA, y0 - p.ph, zi, n);
                // dy: dims (x*4, channel plane, z, y, n)       box (TX*4, PB, kd, TY, 1)
                tma_load_5d(st + p.a_bytes, &tmdy, &full[s], x0 * 4, u.nc * p.PB, zi + p.pd - (p.kd - 1), y0, n);
                if (++s == (uint32_t)p.SA) { s = 0; par ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp, uniform control flow; one elected lane issues) =====================
        const uint32_t idesc = umma_idesc_f16(p.kd * p.NTW, 1, 1);        // both operands MN-major
        const bool leader = elect_one();
        // MN-major, no swizzle: the two 8-voxel K halves of one MMA are 128 B apart (LBO), the 8-channel M / N groups one
        // tile plane apart (SBO)
        const uint64_t a_tmpl = umma_desc(0, 128, (uint32_t)(p.HXa * 16));
        const uint64_t b_tmpl = umma_desc(0, 128, (uint32_t)(p.TX * 16));
        const uint32_t a_hi = (uint32_t)(a_tmpl >> 32), b_hi = (uint32_t)(b_tmpl >> 32);
        const uint32_t a_lo0 = (uint32_t)a_tmpl + (smem_u32(smem) >> 4);
        const uint32_t b_lo0 = a_lo0 - (uint32_t)a_tmpl + (uint32_t)b_tmpl + (p.a_bytes >> 4);
        const uint32_t stage16 = p.stage_bytes >> 4;
        const uint32_t row_a16 = (uint32_t)(p.PA * p.HXa), row_b16 = (uint32_t)(p.kd * p.PB * p.TX);
        const uint32_t acc_cols = (uint32_t)(p.kd * p.NTW);
        const int ksteps = p.TX / 16;
        uint32_t s = 0, par = 0, started = 0;
        for (int L = L0; L < L1; L++) {
            mbar_wait(&full[s], par);
            tc_fence_after();
            uint32_t a_row = a_lo0 + s * stage16, b_row = b_lo0 + s * stage16;
            const int shape = (p.kw == 3 ? 2 : 0) + (ksteps == 2 ? 1 : 0);
            for (int r = 0; r < p.TY; r++) {
                const uint32_t flag = started | (uint32_t)r;          // 0 only for the very first row this CTA contracts
                if (leader) {
                    switch (shape) {
                    case 3: wg_issue_row<3, 2>(tmem_base, acc_cols, a_row, a_hi, b_row, b_hi, idesc, flag); break;
                    case 2: wg_issue_row<3, 1>(tmem_base, acc_cols, a_row, a_hi, b_row, b_hi, idesc, flag); break;
                    case 1: wg_issue_row<1, 2>(tmem_base, acc_cols, a_row, a_hi, b_row, b_hi, idesc, flag); break;
                    default: wg_issue_row<1, 1>(tmem_base, acc_cols, a_row, a_hi, b_row, b_hi, idesc, flag); break;
                    }
                }
                a_row += row_a16; b_row += row_b16;
            }
            started = 1u;
            if (leader) umma_commit(&empty[s]);
            __syncwarp();
            if (++s == (uint32_t)p.SA) { s = 0; par ^= 1; }
        }
        if (leader) umma_commit(done);
        __syncwarp();
    } else {
        // ===================== epilogue: TMEM -> split-K partial =====================
        // accumulator dx, row r = j*CB + c, column zs*NTW + co  <->  tap (dz = kd-1-zs, dy = j, dx), channels (c, co)
        mbar_wait(done, 0);
        tc_fence_after();
        const bool any = L1 > L0;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int j = row / p.CB, c = row % p.CB;
        const int ntaps = p.kd * p.kh * p.kw;
        const bool row_ok = j < p.kh;
        for (int dx = 0; dx < p.kw; dx++) {
            for (int zs = 0; zs < p.kd; zs++) {
                const int tap = ((p.kd - 1 - zs) * p.kh + j) * p.kw + dx;
                for (int cb = 0; cb < p.NTW; cb += 16) {
                    float v[16];
                    if (any) {
                        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((dx * p.kd + zs) * p.NTW + cb), v);
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

// One split-K reduction: S partials of [ntaps][ktot][npad_total] floats -> dw in torch layout.
struct WgReduceJob {
    const float* part; float* dw; const float* dy_unscale;
    int S, ntaps, ktot, npad_total, CB, mchunks0, C0, C1, Co, layout, up_taps, up_co, up_copad;
    int first_block, nblocks, warp_mode;          // batched launch: this job's slice of the grid
};

// deterministic split-K reduction + scatter into the torch parameter layout.  A block of 8 warps owns 128 consecutive
// elements: warp w adds the partials s = w, w + 8, ... (coalesced 512-byte rows of float4), the eight sums meet in shared
// memory and are added in warp order.  (One thread per element walking all S <= 148 partials was latency bound:
// 25 us for the 16 MB of partials of a 32 -> 32 layer; 128-byte rows: 11.7 us.)
E3B_DEVINL void wgrad_reduce_warps(const WgReduceJob& j, unsigned block, unsigned nblocks)
{
    // a block owns 128 consecutive elements (4 per lane: 512-byte rows); npad_total % 4 == 0, so a lane's four elements
    // share (tap, input channel) and differ in the output column only
    __shared__ float4 red[8][32];
    const unsigned total = (unsigned)j.ntaps * j.ktot * j.npad_total;           // < 2^31 (checked by the launch wrapper)
    const float unscale = j.dy_unscale ? __ldg(j.dy_unscale) : 1.f;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (unsigned i0 = block * 128u; i0 < total; i0 += nblocks * 128u) {
        const unsigned i = i0 + lane * 4u;
        const bool inside = i < total;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inside) {
            const float4* q = reinterpret_cast<const float4*>(j.part + i);
            const size_t step = (size_t)total / 4;                       // float4 units between consecutive partials
#pragma unroll 4
            for (int sp = w; sp < j.S; sp += 8) {
                const float4 v = __ldcs(q + (size_t)sp * step);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        }
        red[w][lane] = s;
        __syncthreads();
        if (w == 0 && inside) {
            float4 t = red[0][lane];
#pragma unroll
            for (int k = 1; k < 8; k++) { const float4 r = red[k][lane]; t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w; }
            const unsigned row = i / (unsigned)j.npad_total;
            const int nn0 = (int)(i - row * (unsigned)j.npad_total);
            const int tap = (int)(row / (unsigned)j.ktot), kk = (int)(row - (unsigned)tap * j.ktot);
            bool valid;
            int ci;
            if (kk < j.mchunks0 * j.CB) { valid = kk < j.C0; ci = kk; }
            else { const int k1 = kk - j.mchunks0 * j.CB; valid = k1 < j.C1; ci = j.C0 + k1; }
            const float tv[4] = {t.x * unscale, t.y * unscale, t.z * unscale, t.w * unscale};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int nn = nn0 + e;
                if (!valid) continue;
                if (j.layout == 0) { if (nn < j.Co) j.dw[((size_t)nn * (j.C0 + j.C1) + ci) * j.ntaps + tap] = tv[e]; }
                else if (nn / j.up_copad < j.up_taps && nn % j.up_copad < j.up_co)
                    j.dw[((size_t)ci * j.up_co + nn % j.up_copad) * j.up_taps + nn / j.up_copad] = tv[e];
            }
        }
        __syncthreads();
    }
}

// one thread per element, walking all S partials: for layers whose contraction is split over fewer than 16 CTAs
E3B_DEVINL void wgrad_reduce_flat(const WgReduceJob& j, unsigned block, unsigned nblocks)
{
    const size_t total = (size_t)j.ntaps * j.ktot * j.npad_total;
    const float unscale = j.dy_unscale ? __ldg(j.dy_unscale) : 1.f;
    for (size_t i = (size_t)block * blockDim.x + threadIdx.x; i < total; i += (size_t)nblocks * blockDim.x) {
        const int nn = (int)(i % j.npad_total);
        const int kk = (int)((i / j.npad_total) % j.ktot);
        const int tap = (int)(i / ((size_t)j.npad_total * j.ktot));
        int ci;
        if (kk < j.mchunks0 * j.CB) { if (kk >= j.C0) continue; ci = kk; }
        else { const int k1 = kk - j.mchunks0 * j.CB; if (k1 >= j.C1) continue; ci = j.C0 + k1; }
        if (j.layout == 0) { if (nn >= j.Co) continue; }
        else { if (nn / j.up_copad >= j.up_taps || nn % j.up_copad >= j.up_co) continue; }
        float s = 0.f;
        for (int sp = 0; sp < j.S; sp++) s += j.part[(size_t)sp * total + i];
        s *= unscale;
        if (j.layout == 0) j.dw[((size_t)nn * (j.C0 + j.C1) + ci) * j.ntaps + tap] = s;
        else j.dw[((size_t)ci * j.up_co + nn % j.up_copad) * j.up_taps + nn / j.up_copad] = s;
    }
}

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const WgReduceJob j) { wgrad_reduce_warps(j, blockIdx.x, gridDim.x); }
__global__ void __launch_bounds__(256) wgrad_reduce_flat_kernel(const WgReduceJob j) { wgrad_reduce_flat(j, blockIdx.x, gridDim.x); }

// The reductions of several layers in ONE launch (the weight gradients of a backward pass are only needed by the optimizer:
// a dozen 5-10 us launches, each a ramp-up and a tail between two big kernels, become one).  The jobs travel as a kernel
// parameter, so a CUDA graph captures them by value.
static constexpr int kWgBatch = 16;
struct WgReduceBatch { WgReduceJob j[kWgBatch]; int n; };
__global__ void __launch_bounds__(256) wgrad_reduce_batched_kernel(const __grid_constant__ WgReduceBatch b)
{
    int k = 0;
    while (k + 1 < b.n && (int)blockIdx.x >= b.j[k + 1].first_block) k++;
    const WgReduceJob& j = b.j[k];
    const unsigned local = blockIdx.x - (unsigned)j.first_block;
    if (j.warp_mode) wgrad_reduce_warps(j, local, (unsigned)j.nblocks);
    else wgrad_reduce_flat(j, local, (unsigned)j.nblocks);
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

// QH tensor (N, P, Da, Ha, Wa, 8 halves) seen as 16-byte voxel units of 4 floats: 5D map with the dims in the order the
// shared-memory tile wants them.  zy_order 0: (x*4, plane, y, z, n)  [x tile: rows outermost];
//                                 zy_order 1: (x*4, plane, z, y, n)  [dy tile: the kd planes of a row adjacent].
// (D, H, W): extents of the VIEW starting at `ptr` (a centre-cropped skip tensor is a sub-box of its allocation: everything
// outside the view reads as 0); (Da, Ha, Wa): extents of the allocation (strides).
static int make_qh_map(CUtensorMap* map, const void* ptr, int N, int P, int D, int H, int W, int Da, int Ha, int Wa, int bx,
                       int bp, int by, int bz, int zy_order)
{
    PFN_encodeTiled enc = get_enc();
    if (!enc) return set_error("cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t sy = (cuuint64_t)Wa * 16, sz = (cuuint64_t)Wa * Ha * 16, sp = (cuuint64_t)Wa * Ha * Da * 16, sn = sp * P;
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
    dims[0] = (cuuint64_t)W * 4; box[0] = (cuuint32_t)bx * 4;
    dims[1] = (cuuint64_t)P; strides[0] = sp; box[1] = (cuuint32_t)bp;
    if (zy_order == 0) {
        dims[2] = (cuuint64_t)H; strides[1] = sy; box[2] = (cuuint32_t)by;
        dims[3] = (cuuint64_t)D; strides[2] = sz; box[3] = (cuuint32_t)bz;
    } else {
        dims[2] = (cuuint64_t)D; strides[1] = sz; box[2] = (cuuint32_t)bz;
        dims[3] = (cuuint64_t)H; strides[2] = sy; box[3] = (cuuint32_t)by;
    }
    dims[4] = (cuuint64_t)N; strides[3] = sn; box[4] = 1;
    for (int i = 0; i < 4; i++) if (box[i] > 256) return set_error("wgrad: TMA box dimension > 256");
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, tma_l2_promotion(),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("wgrad: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

static int plan_wgrad(const e3b_wgrad_args* a, WgradParams& p)
{
    memset(&p, 0, sizeof(p));
    p.N = a->N; p.D = a->D; p.H = a->H; p.W = a->W;
    p.kd = a->kd; p.kh = a->kh; p.kw = a->kw; p.pd = a->pd; p.ph = a->ph; p.pw = a->pw;
    p.Do = a->D + 2 * a->pd - a->kd + 1; p.Ho = a->H + 2 * a->ph - a->kh + 1; p.Wo = a->W + 2 * a->pw - a->kw + 1;
    if (p.Do <= 0 || p.Ho <= 0 || p.Wo <= 0) return set_error("wgrad: empty output");
    const int C1 = a->src1 ? a->C1 : 0;
    const int cmax = cpad16(a->C0 > C1 ? a->C0 : C1);
    // channels per M block: with y taps, 32 (four rows stacked in M, three of them taps); without, as many as there are
    p.CB = a->kh > 1 ? (cmax >= 32 ? 32 : 16) : (cmax >= 128 ? 128 : (cmax >= 64 ? 64 : (cmax >= 32 ? 32 : 16)));
    p.RS = 128 / p.CB;
    p.PA = p.CB / 8;
    p.mchunks0 = (cpad16(a->C0) + p.CB - 1) / p.CB;
    p.mchunks = p.mchunks0 + (C1 > 0 ? (cpad16(C1) + p.CB - 1) / p.CB : 0);
    p.ktot = p.mchunks * p.CB;
    // N columns per z-slot: kw accumulators of kd * NTW columns must fit the 512 TMEM columns, one MMA has N = kd * NTW <= 256
    const int npad = cpad16(a->Co);
    int ntw_max = 512 / (a->kw * a->kd); if (ntw_max > 256 / a->kd) ntw_max = 256 / a->kd;
    ntw_max &= ~15;
    if (npad <= ntw_max) p.NTW = npad;
    else { p.NTW = ntw_max; while (p.NTW > 16 && npad % p.NTW) p.NTW -= 16; }
    p.PB = p.NTW / 8;
    p.nchunks_n = (npad + p.NTW - 1) / p.NTW;
    p.npad_total = p.nchunks_n * p.NTW;
    p.units = p.mchunks * p.nchunks_n;
    p.TX = a->W > 16 ? 32 : 16;
    p.HXa = p.TX + a->kw - 1;
    const size_t budget = 227 * 1024 - 2048 - 512;
    int ty = 8;
    if (const char* e = getenv("E3B_WGRAD_TY")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) ty = v; }      // tuning
    for (;; ty >>= 1) {
        p.TY = ty; p.rows_a = ty + p.RS - 1;
        p.a_bytes = (uint32_t)(((size_t)p.rows_a * p.PA * p.HXa * 16 + 127) & ~(size_t)127);
        p.b_bytes = (uint32_t)((size_t)ty * a->kd * p.PB * p.TX * 16);
        p.stage_bytes = p.a_bytes + p.b_bytes;
        p.tx_bytes = (uint32_t)((size_t)p.rows_a * p.PA * p.HXa * 16) + p.b_bytes;
        const bool small_enough = ty == 1 || ty / 2 < p.Ho;      // do not carry rows a small volume does not have
        if ((size_t)2 * p.stage_bytes <= budget && small_enough) break;
        if (ty == 1) return set_error("wgrad: stage does not fit shared memory");
    }
    int sa = (int)(budget / p.stage_bytes); if (sa > 4) sa = 4;
    p.SA = sa;
    p.tiles_x = (p.Wo + p.TX - 1) / p.TX; p.tiles_y = (p.Ho + p.TY - 1) / p.TY;
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

static void fill_reduce_job(const e3b_wgrad_args* a, const WgradParams& p, WgReduceJob& j)
{
    const int ntaps = a->kd * a->kh * a->kw;
    const size_t total = (size_t)ntaps * p.ktot * p.npad_total;
    j.part = p.part; j.dw = a->dw; j.dy_unscale = a->dy_unscale;
    j.S = p.S; j.ntaps = ntaps; j.ktot = p.ktot; j.npad_total = p.npad_total; j.CB = p.CB; j.mchunks0 = p.mchunks0;
    // (source 1's channels start at the next CB boundary after source 0's in the partial row space)
    j.C0 = a->C0; j.C1 = a->src1 ? a->C1 : 0; j.Co = a->Co; j.layout = a->layout; j.up_taps = a->up_taps; j.up_co = a->up_co;
    j.up_copad = cpad8(a->up_co);
    j.first_block = 0;
    j.warp_mode = (p.S >= 16 && total < ((size_t)1 << 31) && p.npad_total % 4 == 0) ? 1 : 0;
    if (j.warp_mode) { int blocks = (int)((total + 127) / 128); if (blocks > 8 * num_sms()) blocks = 8 * num_sms(); j.nblocks = blocks; }
    else { int blocks = (int)((total + 255) / 256); if (blocks > 4 * num_sms()) blocks = 4 * num_sms(); j.nblocks = blocks; }
}

int launch_wgrad_tc(const e3b_wgrad_args* a, cudaStream_t stream)
{
    WgradParams p;
    int rc = plan_wgrad(a, p);
    if (rc) return rc;
    p.part = a->workspace;
    if (((uintptr_t)a->src0 | (uintptr_t)a->src1 | (uintptr_t)a->dy) & 15) return set_error("wgrad: pointers must be 16-byte aligned");
    CUtensorMap mx0, mx1, mdy;
    rc = make_qh_map(&mx0, a->src0, a->N, cpad16(a->C0) / 8, a->D, a->H, a->W, a->D, a->H, a->W, p.HXa, p.PA, p.rows_a, 1, 0);
    if (rc) return rc;
    if (a->src1) {
        // the centre-cropped view of the skip tensor (autocrop, unet.py:303-324): a voxel is a 16-byte unit, any offset is legal
        const uint8_t* v1 = reinterpret_cast<const uint8_t*>(a->src1) +
                            (((size_t)a->off1_d * a->H1 + a->off1_h) * a->W1 + a->off1_w) * 16;
        rc = make_qh_map(&mx1, v1, a->N, cpad16(a->C1) / 8, a->D, a->H, a->W, a->D1, a->H1, a->W1, p.HXa, p.PA, p.rows_a, 1, 0);
        if (rc) return rc;
    } else mx1 = mx0;
    rc = make_qh_map(&mdy, a->dy, a->N, cpad16(a->Co) / 8, p.Do, p.Ho, p.Wo, p.Do, p.Ho, p.Wo, p.TX, p.PB, p.TY, a->kd, 1);
    if (rc) return rc;
    const size_t smem = (size_t)p.SA * p.stage_bytes + 1024 + 1024;
    static bool configured[kMaxDevices] = {false};
    if (!configured[current_device()]) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured[current_device()] = true;
    }
    wgrad_tc_kernel<<<p.units * p.S, kWgThreads, smem, stream>>>(mx0, mx1, mdy, p);
    rc = check_launch("wgrad_tc");
    if (rc) return rc;
    if (a->defer_reduce) return 0;                     // the caller reduces several layers at once (e3b_wgrad_reduce_batched)
    WgReduceJob j;
    fill_reduce_job(a, p, j);
    if (j.warp_mode) wgrad_reduce_kernel<<<j.nblocks, 256, 0, stream>>>(j);
    else wgrad_reduce_flat_kernel<<<j.nblocks, 256, 0, stream>>>(j);
    return check_launch("wgrad_reduce");
}

int launch_wgrad_reduce_batched(const e3b_wgrad_args* args, int n, cudaStream_t stream)
{
    for (int i0 = 0; i0 < n; i0 += kWgBatch) {
        WgReduceBatch b;
        memset(&b, 0, sizeof(b));
        b.n = n - i0 < kWgBatch ? n - i0 : kWgBatch;
        int blocks = 0;
        for (int i = 0; i < b.n; i++) {
            const e3b_wgrad_args* a = args + i0 + i;
            if (!a->workspace || !a->dw) return set_error("wgrad_reduce_batched: null workspace / dw");
            WgradParams p;
            const int rc = plan_wgrad(a, p);
            if (rc) return rc;
            p.part = a->workspace;
            fill_reduce_job(a, p, b.j[i]);
            b.j[i].first_block = blocks;
            blocks += b.j[i].nblocks;
        }
        wgrad_reduce_batched_kernel<<<blocks, 256, 0, stream>>>(b);
        const int rc = check_launch("wgrad_reduce_batched");
        if (rc) return rc;
    }
    return 0;
}

}  // namespace e3b
