// Convolution weight gradient on tcgen05 tensor cores, sm_100a.
//
//   dW[tap][ci][co] = sum over output voxels v of  x[v + tap - pad][ci] * dy[v][co]
//
// (the backward-filter of nn.Conv3d at elektronn3 models/unet.py:131-149; with one tap on
// (x, space-to-depth(dy)) also nn.ConvTranspose3d's, unet.py:152-165).
//
// GEMM view: M = input channels (128 TMEM lanes per CTA), N = output channels, K = voxels.  Both
// operands are read straight from QP tiles as MN-major no-swizzle UMMA operands: 8 consecutive
// x-voxels of a 4-channel plane are exactly one canonical core matrix with K (voxels) 16 bytes apart,
// so one tf32 MMA (K = 8) consumes one 8-voxel row, and the (kh,kw) stencil taps are the same x halo
// tile read through shifted descriptor start addresses.  Each tap owns its own TMEM accumulator
// (TG taps x N columns <= 512).  The contraction over voxels is split over CTAs (split-K); partials go
// to a workspace and a deterministic second kernel reduces them into the torch weight layout.
#include "common.cuh"
#include "kernels.h"

namespace e3b {

static constexpr int kWgThreads = 192;

struct WgradParams {
    int N, D, H, W;              // x extents
    int Do, Ho, Wo;              // dy extents
    int kd, kh, kw, pd, ph, pw;
    int off1_d, off1_h, off1_w;
    int TYW, HX, HYW;            // rows per stage, halo extents
    int tiles_x, tiles_y, total_vt;
    int cq0, cq1;                // planes in source 0 / 1
    int mchunks0, mchunks;       // 128-channel M chunks in source 0 / total
    int NTW, nchunks_n;          // N columns per CTA, number of N chunks
    int TG, nsub;                // taps per CTA (of the kh*kw in-plane taps), subgroups
    int units, S;                // work units, split-K factor
    int stages;
    uint32_t a_plane, a_region, b_plane, b_bytes, stage_bytes;
    int box_planes0, box_planes1, box_planes_dy;   // planes actually moved by each TMA box
    int kpad_total, npad_total;  // padded K (= Cin) and N (= Cout) spaces
    float* part;                 // [S][ntaps][kpad_total][npad_total]
};

struct WgUnit { int kdi, sg, mc, nc; };

E3B_DEVINL WgUnit decode_unit(const WgradParams& p, int u) {
    WgUnit r;
    r.nc = u % p.nchunks_n; u /= p.nchunks_n;
    r.mc = u % p.mchunks; u /= p.mchunks;
    r.sg = u % p.nsub; u /= p.nsub;
    r.kdi = u;
    return r;
}

E3B_DEVINL bool decode_vt(const WgradParams& p, int vt, int kdi, int& n, int& z, int& y0, int& x0, int& zi) {
    int xt = vt % p.tiles_x; vt /= p.tiles_x;
    int yt = vt % p.tiles_y; vt /= p.tiles_y;
    z = vt % p.Do;
    n = vt / p.Do;
    x0 = xt * 8; y0 = yt * p.TYW;
    zi = z + kdi - p.pd;
    return zi >= 0 && zi < p.D;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmx0, const __grid_constant__ CUtensorMap tmx1,
                const __grid_constant__ CUtensorMap tmdy, const WgradParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.stages;
    uint64_t* done = empty + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x / p.S, split = blockIdx.x % p.S;
    const WgUnit u = decode_unit(p, unit);
    const bool src1 = u.mc >= p.mchunks0;
    const int mc_local = src1 ? u.mc - p.mchunks0 : u.mc;
    const int cq_src = src1 ? p.cq1 : p.cq0;
    int planes = cq_src - mc_local * 32; if (planes > 32) planes = 32;   // real planes in this M chunk
    const uint32_t a_bytes = (uint32_t)(src1 ? p.box_planes1 : p.box_planes0) * p.a_plane;
    const uint32_t b_tx = (uint32_t)p.box_planes_dy * p.b_plane;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
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
        if (lane == 0) {
            uint32_t st = 0, ph = 0;
            for (int vt = split; vt < p.total_vt; vt += p.S) {
                int n, z, y0, x0, zi;
                if (!decode_vt(p, vt, u.kdi, n, z, y0, x0, zi)) continue;
                mbar_wait(&empty[st], ph ^ 1);
                uint8_t* sA = smem + (size_t)st * p.stage_bytes;
                uint8_t* sB = sA + p.a_region;
                mbar_arrive_expect_tx(&full[st], a_bytes + b_tx);
                if (!src1)
                    tma_load_5d(sA, &tmx0, &full[st], (x0 - p.pw) * 4, y0 - p.ph, zi, mc_local * 32, n);
                else
                    tma_load_5d(sA, &tmx1, &full[st], (x0 - p.pw + p.off1_w) * 4, y0 - p.ph + p.off1_h,
                                zi + p.off1_d, mc_local * 32, n);
                tma_load_5d(sB, &tmdy, &full[st], x0 * 4, y0, z, u.nc * (p.NTW / 4), n);
                if (++st == (uint32_t)p.stages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(p.NTW, 1, 1);
            uint32_t st = 0, ph = 0;
            bool first = true;
            for (int vt = split; vt < p.total_vt; vt += p.S) {
                int n, z, y0, x0, zi;
                if (!decode_vt(p, vt, u.kdi, n, z, y0, x0, zi)) continue;
                mbar_wait(&full[st], ph);
                tc_fence_after();
                const uint32_t sA = smem_u32(smem + (size_t)st * p.stage_bytes);
                const uint32_t sB = sA + p.a_region;
                for (int tg = 0; tg < p.TG; tg++) {
                    const int tap2 = u.sg * p.TG + tg;          // in-plane tap index
                    const int dy = tap2 / p.kw, dx = tap2 % p.kw;
                    for (int yy = 0; yy < p.TYW; yy++) {
                        const uint64_t ad = umma_desc(sA + (uint32_t)((yy + dy) * p.HX + dx) * 16u, 128, p.a_plane);
                        const uint64_t bd = umma_desc(sB + (uint32_t)(yy * 8) * 16u, 128, p.b_plane);
                        umma_tf32(tmem_base + (uint32_t)(tg * p.NTW), ad, bd, idesc, (first && yy == 0) ? 0u : 1u);
                    }
                }
                first = false;
                umma_commit(&empty[st]);
                if (++st == (uint32_t)p.stages) { st = 0; ph ^= 1; }
            }
            umma_commit(done);
        }
    } else {
        // epilogue: TMEM -> split-K partial
        bool any = false;
        for (int vt = split; vt < p.total_vt && !any; vt += p.S) {
            int n, z, y0, x0, zi;
            any = decode_vt(p, vt, u.kdi, n, z, y0, x0, zi);
        }
        mbar_wait(done, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int row = q * 32 + lane;                       // M row = channel inside the chunk
        const int ntaps = p.kd * p.kh * p.kw;
        const int kbase = (src1 ? p.cq0 * 4 : 0) + mc_local * 128;
        const bool row_ok = row < planes * 4;
        for (int tg = 0; tg < p.TG; tg++) {
            const int tap = u.kdi * (p.kh * p.kw) + u.sg * p.TG + tg;
            for (int cb = 0; cb < p.NTW; cb += 16) {
                float v[16];
                if (any) {
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tg * p.NTW + cb), v);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j++) v[j] = 0.f;
                }
                if (row_ok) {
                    float* o = p.part + (((size_t)split * ntaps + tap) * p.kpad_total + kbase + row) * p.npad_total +
                               u.nc * p.NTW + cb;
#pragma unroll
                    for (int j4 = 0; j4 < 4; j4++)
                        reinterpret_cast<float4*>(o)[j4] = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// deterministic split-K reduction + scatter into the torch parameter layout
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int S, int ntaps,
                                    int kpad_total, int npad_total, int C0, int C0pad, int C1, int Co, int layout,
                                    int up_taps, int up_co, int up_copad)
{
    const size_t total = (size_t)ntaps * kpad_total * npad_total;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int nn = (int)(i % npad_total);
        const int kk = (int)((i / npad_total) % kpad_total);
        const int tap = (int)(i / ((size_t)npad_total * kpad_total));
        int ci;
        if (kk < C0pad) { if (kk >= C0) continue; ci = kk; }
        else { if (kk - C0pad >= C1) continue; ci = C0 + kk - C0pad; }
        float s = 0.f;
        for (int sp = 0; sp < S; sp++) s += part[(size_t)sp * total + i];
        if (layout == 0) {
            if (nn >= Co) continue;
            dw[((size_t)nn * (C0 + C1) + ci) * ntaps + tap] = s;
        } else {
            const int t = nn / up_copad, co = nn % up_copad;
            if (t >= up_taps || co >= up_co) continue;
            dw[((size_t)ci * up_co + co) * up_taps + t] = s;
        }
    }
}

static int plan_wgrad(const e3b_wgrad_args* a, WgradParams& p)
{
    memset(&p, 0, sizeof(p));
    p.N = a->N; p.D = a->D; p.H = a->H; p.W = a->W;
    p.kd = a->kd; p.kh = a->kh; p.kw = a->kw; p.pd = a->pd; p.ph = a->ph; p.pw = a->pw;
    p.Do = a->D + 2 * a->pd - a->kd + 1; p.Ho = a->H + 2 * a->ph - a->kh + 1; p.Wo = a->W + 2 * a->pw - a->kw + 1;
    p.off1_d = a->off1_d; p.off1_h = a->off1_h; p.off1_w = a->off1_w;
    p.cq0 = cpad8(a->C0) / 4;
    p.cq1 = a->src1 ? cpad8(a->C1) / 4 : 0;
    p.mchunks0 = (p.cq0 + 31) / 32;
    p.mchunks = p.mchunks0 + (p.cq1 + 31) / 32;
    p.kpad_total = (p.cq0 + p.cq1) * 4;
    const int copad8 = cpad8(a->Co);          // dy tensor planes
    p.npad_total = cpad16(a->Co);
    p.NTW = conv_ntile_width(p.npad_total);
    if (p.NTW <= 0) return set_error("wgrad: unsupported output width %d", p.npad_total);
    (void)copad8;
    p.nchunks_n = p.npad_total / p.NTW;
    const int inplane = a->kh * a->kw;
    int tg = inplane;
    while (tg > 1 && tg * p.NTW > 512) tg /= 3;
    if (tg * p.NTW > 512) return set_error("wgrad: N tile too wide");
    p.TG = tg; p.nsub = inplane / tg;
    p.units = a->kd * p.nsub * p.mchunks * p.nchunks_n;
    p.HX = 8 + a->kw - 1;
    const size_t budget = 227 * 1024 - 1024 - 256;
    int tyw = 16;
    for (;; tyw >>= 1) {
        p.TYW = tyw; p.HYW = tyw + a->kh - 1;
        p.a_plane = (uint32_t)(p.HX * p.HYW * 16);
        p.a_region = 32u * p.a_plane;
        p.b_plane = (uint32_t)(8 * tyw * 16);
        p.b_bytes = p.b_plane * (uint32_t)(p.NTW / 4);
        p.stage_bytes = p.a_region + p.b_bytes;
        int st = (int)(budget / p.stage_bytes);
        if (st >= 3 || (st >= 2 && tyw == 1)) { p.stages = st > 4 ? 4 : st; break; }
        if (tyw == 1) return set_error("wgrad: stage does not fit shared memory");
    }
    while (p.TYW > 1 && p.TYW / 2 >= p.Ho) { p.TYW >>= 1; }   // do not carry rows a small volume does not have
    if (p.TYW != tyw) {
        p.HYW = p.TYW + a->kh - 1;
        p.a_plane = (uint32_t)(p.HX * p.HYW * 16); p.a_region = 32u * p.a_plane;
        p.b_plane = (uint32_t)(8 * p.TYW * 16); p.b_bytes = p.b_plane * (uint32_t)(p.NTW / 4);
        p.stage_bytes = p.a_region + p.b_bytes;
        int st = (int)(budget / p.stage_bytes); p.stages = st > 4 ? 4 : st;
    }
    p.box_planes0 = p.cq0 < 32 ? p.cq0 : 32;
    p.box_planes1 = p.cq1 < 32 ? p.cq1 : 32;
    { const int cqdy = cpad8(a->Co) / 4; p.box_planes_dy = cqdy < p.NTW / 4 ? cqdy : p.NTW / 4; }
    p.tiles_x = (p.Wo + 7) / 8; p.tiles_y = (p.Ho + p.TYW - 1) / p.TYW;
    p.total_vt = p.tiles_x * p.tiles_y * p.Do * a->N;
    int S = (2 * num_sms()) / p.units; if (S < 1) S = 1;
    if (S > p.total_vt) S = p.total_vt;
    if (S > 64) S = 64;
    p.S = S;
    return 0;
}

int64_t wgrad_workspace_floats(const e3b_wgrad_args* a)
{
    WgradParams p;
    if (plan_wgrad(a, p)) return -1;
    return (int64_t)p.S * a->kd * a->kh * a->kw * p.kpad_total * p.npad_total;
}

int launch_wgrad_tc(const e3b_wgrad_args* a, cudaStream_t stream)
{
    WgradParams p;
    int rc = plan_wgrad(a, p);
    if (rc) return rc;
    p.part = a->workspace;
    CUtensorMap mx0, mx1, mdy;
    rc = make_qp_tensor_map(&mx0, a->src0, a->N, p.cq0, a->D, a->H, a->W, p.HX, p.HYW, 1, p.box_planes0);
    if (rc) return rc;
    if (a->src1) {
        rc = make_qp_tensor_map(&mx1, a->src1, a->N, p.cq1, a->D1, a->H1, a->W1, p.HX, p.HYW, 1, p.box_planes1);
        if (rc) return rc;
    } else mx1 = mx0;
    rc = make_qp_tensor_map(&mdy, a->dy, a->N, cpad8(a->Co) / 4, p.Do, p.Ho, p.Wo, 8, p.TYW, 1, p.box_planes_dy);
    if (rc) return rc;
    const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    wgrad_tc_kernel<<<p.units * p.S, kWgThreads, smem, stream>>>(mx0, mx1, mdy, p);
    rc = check_launch("wgrad_tc");
    if (rc) return rc;
    const int ntaps = a->kd * a->kh * a->kw;
    const size_t total = (size_t)ntaps * p.kpad_total * p.npad_total;
    int blocks = (int)((total + 255) / 256); if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
    wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(p.part, a->dw, p.S, ntaps, p.kpad_total, p.npad_total, a->C0,
                                                    cpad8(a->C0), a->src1 ? a->C1 : 0, a->Co, a->layout, a->up_taps,
                                                    a->up_co, cpad8(a->up_co));
    return check_launch("wgrad_reduce");
}

}  // namespace e3b
