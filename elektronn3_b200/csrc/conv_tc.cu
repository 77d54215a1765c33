// Implicit-GEMM 3D convolution (forward, dgrad, and the k=s=2 transposed conv as a 1-tap GEMM with a
// pixel-shuffle scatter epilogue) on tcgen05 tensor cores, sm_100a.
//
// Replaces the torch/cuDNN calls behind elektronn3 `conv3` (models/unet.py:131-149), `conv1`
// (:178-180) and `upconv2` (:152-165).
//
// Design ("halo-tile implicit GEMM"):
//  * the operand is a QH tensor (fp16, 8 channels per 16-byte voxel unit, see common.cuh); the output an
//    fp32 QP tensor (or QH again when the epilogue already produces the next layer's operand).  A CTA tile
//    is 8(x) x 16(y) x TZ(z) output voxels = TZ accumulators of M=128 rows x N columns in TMEM.
//  * per 16 input channels (two 16-byte planes = the K of one kind::f16 MMA) ONE TMA box load brings the
//    (8+kw-1) x (16+kh-1) x (TZ+kd-1) halo tile of those channels into shared memory (zero fill outside the volume == conv zero padding).  Because
//    one voxel is a 16-byte quad in the no-swizzle canonical UMMA layout, every stencil tap is the
//    same tile read through a descriptor whose start address is shifted by
//    ((dz*HY + dy)*HX + dx) * 16 B: 27 taps x TZ planes of MMAs are issued per loaded tile, so L2->smem
//    traffic is ~1.4x the activation size instead of 27x.
//  * weights are pre-packed into the exact smem image (K-major, no swizzle) and streamed with 1D bulk
//    copies in tap groups through their own mbarrier ring.
//  * warp roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2..5 and 6..9 two epilogue
//    warpgroups on alternate 16-column groups (tcgen05.ld -> +bias -> [ReLU] -> per-channel sum /
//    sum-of-squares for the following Group/BatchNorm -> coalesced 16-byte stores; the transposed conv's
//    scatter pairs the two x taps of a coarse voxel into one 32-byte store).  Two TMEM accumulator sets ping-pong so the epilogue
//    of tile i overlaps the MMAs of tile i+1.  CTAs are persistent over a static tile schedule.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace e3b {

static constexpr int kTX = 8, kTY = 16;
static constexpr int kThreads = 320;      // warp 0 producer, warp 1 MMA issuer, warps 2..9: two epilogue warpgroups
static constexpr int kEpiThreads = 256;

struct ConvTcParams {
    // geometry
    int N, Do, Ho, Wo;           // output extents (the GEMM M space)
    int kd, kh, kw;              // taps per dim (1 or 3)
    int pd, ph, pw;              // zero padding
    int TZ;                      // output planes per tile
    int HX, HY, HZ;              // halo tile extents
    int tiles_x, tiles_y, tiles_z, n_ntiles, total_tiles;
    int chunks0, chunks1;        // 16-channel K chunks from source 0 / source 1
    int off1_d, off1_h, off1_w;  // crop offset into source 1
    int NT;                      // columns per N tile (multiple of 16, <= 256)
    int TG;                      // taps per weight stage
    int SA, SB;                  // pipeline depths
    uint32_t a_stage_bytes, b_stage_bytes;
    // epilogue
    const float* bias;           // [n_bias] or null
    int n_bias;
    float* dst0; int cq0;        // channel planes [0, cq0) of the N space go to dst0 ...
    float* dst1;                 // ... the rest to dst1 (dgrad of a virtual-concat conv)
    int cq0_alloc, cq1_alloc;    // planes allocated in dst0 / dst1
    int relu;
    int debug;                   // accumulate role timings into g_conv_dbg
    int half_out;                // store the output as a QH (fp16) operand tensor: it feeds the next MMA
    const float* out_scale;      // optional device scalar multiplied into the accumulator (gradient un-scaling)
    const float* w_unscale;      // optional second one (un-scaling of power-of-two scaled weights)
    double* stats;               // [N][Cstat][2] sum / sumsq (fp64 atomics) or null
    int Cstat;
    int scatter;                 // 1: k=s transposed conv, column n = tap*Cup + co, dst is the fine grid
    int sd, sh, sw;              // scatter strides
    int Cup;                     // padded channels per tap in scatter mode
    int Ds, Hs, Ws;              // (scatter) cropped fine-grid output extents
    const float* wpk;
};

static constexpr int kStatSlots = 256;

// Optional cycle accounting of the pipeline roles (scripts/conv_pipeline_debug.py): enabled per launch.
__device__ unsigned long long g_conv_dbg[16];
#define DBG_T0(var) long long var = 0; if (p.debug) var = clock64()
#define DBG_ACC(slot, var) if (p.debug) dbg[slot] += (unsigned long long)(clock64() - var)

// per-CTA statistics accumulators -> global [N][Cstat][2] (fp64 atomics: a few hundred per CTA, not per tile)
E3B_DEVINL void flush_stats(const ConvTcParams& p, double* cta_stats, int n, int nt, int etid) {
    const int nslots = p.scatter ? (p.Cup < kStatSlots ? p.Cup : kStatSlots) : (p.NT < kStatSlots ? p.NT : kStatSlots);
    for (int i = etid; i < nslots * 2; i += kEpiThreads) {
        const int si = i >> 1;
        const int ch = p.scatter ? si : nt * p.NT + si;
        const double v = cta_stats[i];
        cta_stats[i] = 0.0;
        if (ch < p.Cstat && v != 0.0) atomicAdd(p.stats + ((size_t)n * p.Cstat + ch) * 2 + (i & 1), v);
    }
}

// Issue the TZ plane MMAs of one stencil tap (fully unrolled: two uniform adds per MMA).
template <int TZ>
E3B_DEVINL void issue_tap_planes(uint32_t acc, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t accum, uint32_t plane_step,
                                 uint32_t nt)
{
#pragma unroll
    for (int pl = 0; pl < TZ; pl++) umma_f16(acc + (uint32_t)pl * nt, ad + (uint64_t)((uint32_t)pl * plane_step), bd, idesc, accum);
}

E3B_DEVINL void issue_tap(int tz, uint32_t acc, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t accum, uint32_t plane_step,
                          uint32_t nt)
{
    switch (tz) {
    case 1: issue_tap_planes<1>(acc, ad, bd, idesc, accum, plane_step, nt); break;
    case 2: issue_tap_planes<2>(acc, ad, bd, idesc, accum, plane_step, nt); break;
    case 3: issue_tap_planes<3>(acc, ad, bd, idesc, accum, plane_step, nt); break;
    case 4: issue_tap_planes<4>(acc, ad, bd, idesc, accum, plane_step, nt); break;
    case 5: issue_tap_planes<5>(acc, ad, bd, idesc, accum, plane_step, nt); break;
    case 6: issue_tap_planes<6>(acc, ad, bd, idesc, accum, plane_step, nt); break;
    case 7: issue_tap_planes<7>(acc, ad, bd, idesc, accum, plane_step, nt); break;
    default: issue_tap_planes<8>(acc, ad, bd, idesc, accum, plane_step, nt); break;
    }
}

E3B_DEVINL void decode_tile(const ConvTcParams& p, int t, int& nt, int& n, int& z0, int& y0, int& x0) {
    int xt = t % p.tiles_x; t /= p.tiles_x;
    int yt = t % p.tiles_y; t /= p.tiles_y;
    int zt = t % p.tiles_z; t /= p.tiles_z;
    nt = t % p.n_ntiles;
    n = t / p.n_ntiles;
    x0 = xt * kTX; y0 = yt * kTY; z0 = zt * p.TZ;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
               const ConvTcParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    // carve: [A stages][B stages][barriers]
    uint8_t* a_base = smem;
    uint8_t* b_base = smem + (size_t)p.SA * p.a_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (size_t)p.SB * p.b_stage_bytes);
    uint64_t* a_full = bars;              // [SA]
    uint64_t* a_empty = a_full + p.SA;    // [SA]
    uint64_t* b_full = a_empty + p.SA;    // [SB]
    uint64_t* b_empty = b_full + p.SB;    // [SB]
    uint64_t* acc_full = b_empty + p.SB;  // [2]
    uint64_t* acc_empty = acc_full + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    __shared__ double cta_stats[kStatSlots * 2];
    __shared__ uint32_t tap_off[27];         // halo-tile offset of every stencil tap, in 16-byte (voxel) units
    if (threadIdx.x < 27) {
        const int tp = threadIdx.x;
        tap_off[tp] = (uint32_t)(((tp / (p.kh * p.kw)) * p.HY + (tp / p.kw) % p.kh) * p.HX + tp % p.kw);
    }
    for (int i = threadIdx.x; i < kStatSlots * 2; i += blockDim.x) cta_stats[i] = 0.0;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ntaps = p.kd * p.kh * p.kw;
    const int ngroups = ntaps / p.TG;
    const int nchunks = p.chunks0 + p.chunks1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.SA; i++) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < p.SB; i++) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
        fence_barrier_init();
        tma_prefetch_desc(&tmap0);
        if (p.chunks1) tma_prefetch_desc(&tmap1);
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            unsigned long long dbg[16] = {0};
            DBG_T0(t_all);
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                int nt, n, z0, y0, x0;
                decode_tile(p, t, nt, n, z0, y0, x0);
                for (int c = 0; c < nchunks; c++) {
                    DBG_T0(t0);
                    mbar_wait(&a_empty[sa], pa ^ 1);
                    DBG_ACC(0, t0);
                    mbar_arrive_expect_tx(&a_full[sa], p.a_stage_bytes);
                    if (c < p.chunks0)
                        tma_load_5d(a_base + (size_t)sa * p.a_stage_bytes, &tmap0, &a_full[sa],
                                    (x0 - p.pw) * 4, y0 - p.ph, z0 - p.pd, c * 2, n);
                    else
                        tma_load_5d(a_base + (size_t)sa * p.a_stage_bytes, &tmap1, &a_full[sa],
                                    (x0 - p.pw) * 4, y0 - p.ph, z0 - p.pd, (c - p.chunks0) * 2, n);
                    if (++sa == (uint32_t)p.SA) { sa = 0; pa ^= 1; }
                    for (int g = 0; g < ngroups; g++) {
                        DBG_T0(t1);
                        mbar_wait(&b_empty[sb], pb ^ 1);
                        DBG_ACC(1, t1);
                        mbar_arrive_expect_tx(&b_full[sb], p.b_stage_bytes);
                        const float* src = p.wpk + ((size_t)(nt * nchunks + c) * ntaps + (size_t)g * p.TG) *
                                                       (size_t)(2 * p.NT * 4);
                        bulk_load_1d(b_base + (size_t)sb * p.b_stage_bytes, src, p.b_stage_bytes, &b_full[sb]);
                        if (++sb == (uint32_t)p.SB) { sb = 0; pb ^= 1; }
                    }
                }
            }
            DBG_ACC(2, t_all);
            if (p.debug) for (int i = 0; i < 3; i++) atomicAdd(&g_conv_dbg[i], dbg[i]);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp runs the loop (warp-uniform control flow keeps the descriptor arithmetic on the
        // uniform datapath); one elected lane issues the MMAs and the commits.
        const uint32_t idesc = umma_idesc_f16(p.NT, 0, 0);
        // descriptor templates with a zero address field; one voxel == 16 B == one address unit
        const uint64_t a_tmpl = umma_desc(0, (uint32_t)(p.HX * p.HY * p.HZ * 16), (uint32_t)(p.HX * 16));
        const uint64_t b_tmpl = umma_desc(0, (uint32_t)(p.NT * 16), 128);
        const uint32_t plane_step = (uint32_t)(p.HY * p.HX);
        const uint32_t tap_step_b = (uint32_t)(2 * p.NT);
        const bool leader = elect_one();
        unsigned long long dbg[16] = {0};
        DBG_T0(t_all);
        uint32_t sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, it++) {
            const uint32_t buf = it & 1, use = it >> 1;
            DBG_T0(t0);
            mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
            DBG_ACC(3, t0);
            tc_fence_after();
            const uint32_t acc = tmem_base + buf * 256;
            for (int c = 0; c < nchunks; c++) {
                DBG_T0(t1);
                mbar_wait(&a_full[sa], pa);
                DBG_ACC(4, t1);
                tc_fence_after();
                const uint32_t a16 = smem_u32(a_base + (size_t)sa * p.a_stage_bytes) >> 4;
                int tap = 0;
                for (int g = 0; g < ngroups; g++) {
                    DBG_T0(t2);
                    mbar_wait(&b_full[sb], pb);
                    DBG_ACC(5, t2);
                    tc_fence_after();
                    DBG_T0(t3);
                    const uint32_t b16 = smem_u32(b_base + (size_t)sb * p.b_stage_bytes) >> 4;
                    if (leader) {
                        uint64_t bd = b_tmpl + b16;
                        const uint64_t a_stage = a_tmpl + a16;
                        for (int tg = 0; tg < p.TG; tg++, tap++) {
                            issue_tap(p.TZ, acc, a_stage + tap_off[tap], bd, idesc, (c | tap) ? 1u : 0u, plane_step, (uint32_t)p.NT);
                            bd += tap_step_b;
                        }
                        umma_commit(&b_empty[sb]);
                    }
                    __syncwarp();
                    DBG_ACC(6, t3);
                    if (++sb == (uint32_t)p.SB) { sb = 0; pb ^= 1; }
                }
                if (leader) umma_commit(&a_empty[sa]);
                __syncwarp();
                if (++sa == (uint32_t)p.SA) { sa = 0; pa ^= 1; }
            }
            if (leader) umma_commit(&acc_full[buf]);
            __syncwarp();
        }
        DBG_ACC(7, t_all);
        if (p.debug && leader) for (int i = 3; i < 8; i++) atomicAdd(&g_conv_dbg[i], dbg[i]);
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // GEMM row inside the tile
        const int ry = row >> 3, rx = row & 7;
        const int etid = threadIdx.x - 64;      // 0..255
        const int eg = (warp - 2) >> 2;          // epilogue warpgroup: the two groups take alternate 16-column groups
        // butterfly transpose-reduce leaves column (bit-reversed low nibble of the lane) in lanes 0..15
        const int bcol = ((lane & 1) << 3) | ((lane & 2) << 1) | ((lane & 4) >> 1) | ((lane & 8) >> 3);
        int cur_n = -1, cur_nt = -1;
        uint32_t it = 0;
        unsigned long long dbg[16] = {0};
        DBG_T0(t_all);
        const float oscale = (p.out_scale ? __ldg(p.out_scale) : 1.f) * (p.w_unscale ? __ldg(p.w_unscale) : 1.f);
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, it++) {
            const uint32_t buf = it & 1, use = it >> 1;
            int nt, n, z0, y0, x0;
            decode_tile(p, t, nt, n, z0, y0, x0);
            if (p.stats && (n != cur_n || (!p.scatter && nt != cur_nt))) {
                // the per-CTA statistics accumulators belong to one (sample, N tile): flush on change
                named_bar_sync(1, kEpiThreads);
                if (cur_n >= 0) flush_stats(p, cta_stats, cur_n, cur_nt, etid);
                named_bar_sync(1, kEpiThreads);
                cur_n = n; cur_nt = nt;
            }
            DBG_T0(t0);
            mbar_wait(&acc_full[buf], use & 1);
            DBG_ACC(8, t0);
            tc_fence_after();
            const int y = y0 + ry, x = x0 + rx;
            const bool valid = (y < p.Ho) && (x < p.Wo);
            const int npl = (p.Do - z0) < p.TZ ? (p.Do - z0) : p.TZ;
            // Column groups of 16.  In scatter mode the columns of one tile are [tap][channel]: the groups are
            // visited channel-major so that the statistics of a channel group are reduced once for all taps.
            const int cspan = (p.scatter && p.Cup < p.NT) ? p.Cup : p.NT;      // columns per tap inside this tile
            const int ntp = p.NT / cspan;                                      // taps inside this tile
            // Transposed conv with stride 2 along x: the taps (.., tk = 0) and (.., tk = 1) of a coarse voxel are adjacent
            // fine voxels.  Handling them together turns two half-sector stores (16 B at a 32 B stride) into one 32-byte
            // store per lane -- the scatter epilogue is bound by L2 write transactions, not by bytes.
            const bool pair_x = p.scatter && p.sw == 2 && (ntp & 1) == 0;
            const bool wide_ok = (p.Ws & 1) == 0;          // every fine row starts 32-byte aligned
            for (int cg = eg * 16; pair_x && cg < cspan; cg += 32) {
                float s[16], ss[16];
#pragma unroll
                for (int j = 0; j < 16; j++) { s[j] = 0.f; ss[j] = 0.f; }
                const int co = (nt * p.NT + cg) % p.Cup;
                float bias_v[16];
#pragma unroll
                for (int j = 0; j < 16; j++) bias_v[j] = (p.bias && co + j < p.n_bias) ? __ldg(p.bias + co + j) : 0.f;
                for (int tp = 0; tp < ntp; tp += 2) {
                    const int cb = tp * cspan + cg;
                    const int tapi = (nt * p.NT + cb) / p.Cup;          // even: tk = 0; its partner tapi + 1 has tk = 1
                    const int ti = tapi / (p.sh * p.sw), tj = (tapi / p.sw) % p.sh;
                    for (int pl = 0; pl < npl; pl++) {
                        const int z = z0 + pl;
                        uint32_t r[32];
                        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256 + (uint32_t)(pl * p.NT + cb);
                        tmem_ld16_raw(taddr, r);
                        tmem_ld16_raw(taddr + (uint32_t)cspan, r + 16);
                        tmem_ld_wait32(r);
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            v[j] = fmaf(__uint_as_float(r[j]), oscale, bias_v[j & 15]);
                            if (p.relu) v[j] = fmaxf(v[j], 0.f);
                        }
                        const int fz = z * p.sd + ti, fy = y * p.sh + tj, fx = x * 2;
                        const bool sv0 = valid && fz < p.Ds && fy < p.Hs && fx < p.Ws;
                        const bool sv1 = sv0 && fx + 1 < p.Ws;
                        if (sv0) {
                            if (p.half_out) {
                                // operand tensor (QH): 16 channels = the 16-byte units of planes co/8 and co/8 + 1
#pragma unroll
                                for (int j8 = 0; j8 < 2; j8++) {
                                    const int hpl = (co >> 3) + j8;
                                    if (hpl < p.cq0_alloc) {
                                        const uint2 a0 = pack_half4(v[j8 * 8], v[j8 * 8 + 1], v[j8 * 8 + 2], v[j8 * 8 + 3]);
                                        const uint2 a1 = pack_half4(v[j8 * 8 + 4], v[j8 * 8 + 5], v[j8 * 8 + 6], v[j8 * 8 + 7]);
                                        const uint2 b0 = pack_half4(v[16 + j8 * 8], v[16 + j8 * 8 + 1], v[16 + j8 * 8 + 2], v[16 + j8 * 8 + 3]);
                                        const uint2 b1 = pack_half4(v[16 + j8 * 8 + 4], v[16 + j8 * 8 + 5], v[16 + j8 * 8 + 6], v[16 + j8 * 8 + 7]);
                                        uint4* o = reinterpret_cast<uint4*>(p.dst0) +
                                                   (((((size_t)n * p.cq0_alloc + hpl) * p.Ds + fz) * p.Hs + fy) * (size_t)p.Ws + fx);
                                        const uint4 u0 = make_uint4(a0.x, a0.y, a1.x, a1.y), u1 = make_uint4(b0.x, b0.y, b1.x, b1.y);
                                        if (sv1 && wide_ok) st_global_256(o, u0, u1);
                                        else { o[0] = u0; if (sv1) o[1] = u1; }
                                    }
                                }
                            } else {
#pragma unroll
                                for (int j4 = 0; j4 < 4; j4++) {
                                    const int cq = (co >> 2) + j4;
                                    if (cq < p.cq0_alloc) {
                                        float4* o = reinterpret_cast<float4*>(p.dst0) +
                                                    (((((size_t)n * p.cq0_alloc + cq) * p.Ds + fz) * p.Hs + fy) * (size_t)p.Ws + fx);
                                        const uint4 u0 = make_uint4(__float_as_uint(v[j4 * 4]), __float_as_uint(v[j4 * 4 + 1]),
                                                                    __float_as_uint(v[j4 * 4 + 2]), __float_as_uint(v[j4 * 4 + 3]));
                                        const uint4 u1 = make_uint4(__float_as_uint(v[16 + j4 * 4]), __float_as_uint(v[16 + j4 * 4 + 1]),
                                                                    __float_as_uint(v[16 + j4 * 4 + 2]), __float_as_uint(v[16 + j4 * 4 + 3]));
                                        if (sv1 && wide_ok) st_global_256(o, u0, u1);
                                        else { *reinterpret_cast<uint4*>(o) = u0; if (sv1) *reinterpret_cast<uint4*>(o + 1) = u1; }
                                    }
                                }
                            }
                        }
                        if (p.stats && sv0) {
#pragma unroll
                            for (int j = 0; j < 16; j++) { s[j] += v[j]; ss[j] = fmaf(v[j], v[j], ss[j]); }
                            if (sv1) {
#pragma unroll
                                for (int j = 0; j < 16; j++) { s[j] += v[16 + j]; ss[j] = fmaf(v[16 + j], v[16 + j], ss[j]); }
                            }
                        }
                    }
                }
                if (p.stats) {
                    // per-channel sum / sumsq over the warp's 32 rows x npl planes x taps: 16-value butterfly
#pragma unroll
                    for (int step = 0; step < 4; step++) {
                        const int keepn = 8 >> step;
                        const int bit = 1 << step;
                        const bool upper = (lane & bit) != 0;
#pragma unroll
                        for (int j = 0; j < keepn; j++) {
                            float send_s = upper ? s[j] : s[j + keepn];
                            float send_q = upper ? ss[j] : ss[j + keepn];
                            float keep_s = upper ? s[j + keepn] : s[j];
                            float keep_q = upper ? ss[j + keepn] : ss[j];
                            s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, bit);
                            ss[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, bit);
                        }
                    }
                    const float fs = s[0] + __shfl_xor_sync(0xffffffffu, s[0], 16);
                    const float fq = ss[0] + __shfl_xor_sync(0xffffffffu, ss[0], 16);
                    if (lane < 16) {
                        const int si = co + bcol;
                        if (si < kStatSlots) {
                            atomicAdd(&cta_stats[si * 2], (double)fs);
                            atomicAdd(&cta_stats[si * 2 + 1], (double)fq);
                        } else if (si < p.Cstat) {
                            atomicAdd(p.stats + ((size_t)n * p.Cstat + si) * 2, (double)fs);
                            atomicAdd(p.stats + ((size_t)n * p.Cstat + si) * 2 + 1, (double)fq);
                        }
                    }
                }
            }
            for (int cg = eg * 16; !pair_x && cg < cspan; cg += 32) {
                float s[16], ss[16];
#pragma unroll
                for (int j = 0; j < 16; j++) { s[j] = 0.f; ss[j] = 0.f; }
                const int co = p.scatter ? ((nt * p.NT + cg) % p.Cup) : (nt * p.NT + cg);   // first channel of the group
                float bias_v[16];
#pragma unroll
                for (int j = 0; j < 16; j++) bias_v[j] = (p.bias && co + j < p.n_bias) ? __ldg(p.bias + co + j) : 0.f;
                for (int tp = 0; tp < ntp; tp++) {
                    const int cb = tp * cspan + cg;
                    const int ncol = nt * p.NT + cb;     // first global N column of this group
                    // scatter geometry of this column group (one tap per 16-column group: Cup % 16 == 0)
                    int ti = 0, tj = 0, tk = 0;
                    if (p.scatter) {
                        const int tapi = ncol / p.Cup;
                        ti = tapi / (p.sh * p.sw); tj = (tapi / p.sw) % p.sh; tk = tapi % p.sw;
                    }
                    for (int pl = 0; pl < npl; pl++) {
                        const int z = z0 + pl;
                        float v[16];
                        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256 + (uint32_t)(pl * p.NT + cb), v);
#pragma unroll
                        for (int j = 0; j < 16; j++) v[j] = fmaf(v[j], oscale, bias_v[j]);
                        if (p.relu) {
#pragma unroll
                            for (int j = 0; j < 16; j++) v[j] = fmaxf(v[j], 0.f);
                        }
                        bool sv = valid;
                        if (p.half_out) {
                            // operand tensor (QH): 16 columns = two 16-byte units (planes ncol/8 and ncol/8 + 1)
                            int fz = z, fy = y, fx = x, hd = p.Do, hh = p.Ho, hw = p.Wo, hp = ncol >> 3;
                            if (p.scatter) {
                                fz = z * p.sd + ti; fy = y * p.sh + tj; fx = x * p.sw + tk;
                                hd = p.Ds; hh = p.Hs; hw = p.Ws; hp = co >> 3;
                                sv = valid && fz < p.Ds && fy < p.Hs && fx < p.Ws;
                            }
                            if (sv) {
#pragma unroll
                                for (int j8 = 0; j8 < 2; j8++) {
                                    const int hpl = hp + j8;
                                    if (hpl < p.cq0_alloc) {
                                        const uint2 lo = pack_half4(v[j8 * 8], v[j8 * 8 + 1], v[j8 * 8 + 2], v[j8 * 8 + 3]);
                                        const uint2 hi = pack_half4(v[j8 * 8 + 4], v[j8 * 8 + 5], v[j8 * 8 + 6], v[j8 * 8 + 7]);
                                        size_t o = ((((size_t)n * p.cq0_alloc + hpl) * hd + fz) * hh + fy) * (size_t)hw + fx;
                                        reinterpret_cast<uint4*>(p.dst0)[o] = make_uint4(lo.x, lo.y, hi.x, hi.y);
                                    }
                                }
                            }
                        } else if (!p.scatter) {
                            if (valid) {
#pragma unroll
                                for (int j4 = 0; j4 < 4; j4++) {
                                    int cq = (ncol >> 2) + j4;
                                    float* base; int cqa;
                                    if (cq < p.cq0) { base = p.dst0; cqa = p.cq0_alloc; }
                                    else { base = p.dst1; cq -= p.cq0; cqa = p.cq1_alloc; }
                                    if (base != nullptr && cq < cqa) {
                                        size_t o = ((((size_t)n * cqa + cq) * p.Do + z) * p.Ho + y) * (size_t)p.Wo + x;
                                        reinterpret_cast<float4*>(base)[o] =
                                            make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                                    }
                                }
                            }
                        } else {
                            // column = tap * Cup + co ; fine voxel = (z*sd+i, y*sh+j, x*sw+k)
                            const int fz = z * p.sd + ti, fy = y * p.sh + tj, fx = x * p.sw + tk;
                            sv = valid && fz < p.Ds && fy < p.Hs && fx < p.Ws;
                            if (sv) {
#pragma unroll
                                for (int j4 = 0; j4 < 4; j4++) {
                                    int cq = (co >> 2) + j4;
                                    if (cq < p.cq0_alloc) {
                                        size_t o = ((((size_t)n * p.cq0_alloc + cq) * p.Ds + fz) * p.Hs + fy) * (size_t)p.Ws + fx;
                                        reinterpret_cast<float4*>(p.dst0)[o] =
                                            make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                                    }
                                }
                            }
                        }
                        if (p.stats && sv) {
#pragma unroll
                            for (int j = 0; j < 16; j++) { s[j] += v[j]; ss[j] = fmaf(v[j], v[j], ss[j]); }
                        }
                    }
                }
                if (p.stats) {
                    // per-channel sum / sumsq over the warp's 32 rows x npl planes (x taps): 16-value butterfly
#pragma unroll
                    for (int step = 0; step < 4; step++) {
                        const int keepn = 8 >> step;          // 8,4,2,1 values kept
                        const int bit = 1 << step;            // exchange partner lane bit
                        const bool upper = (lane & bit) != 0;
#pragma unroll
                        for (int j = 0; j < keepn; j++) {
                            float send_s = upper ? s[j] : s[j + keepn];
                            float send_q = upper ? ss[j] : ss[j + keepn];
                            float keep_s = upper ? s[j + keepn] : s[j];
                            float keep_q = upper ? ss[j + keepn] : ss[j];
                            s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, bit);
                            ss[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, bit);
                        }
                    }
                    const float fs = s[0] + __shfl_xor_sync(0xffffffffu, s[0], 16);
                    const float fq = ss[0] + __shfl_xor_sync(0xffffffffu, ss[0], 16);
                    if (lane < 16) {
                        const int si = p.scatter ? (co + bcol) : (cg + bcol);     // accumulator slot
                        if (si < kStatSlots) {
                            atomicAdd(&cta_stats[si * 2], (double)fs);
                            atomicAdd(&cta_stats[si * 2 + 1], (double)fq);
                        } else {
                            const int ch = p.scatter ? si : nt * p.NT + cg + bcol;
                            if (ch < p.Cstat) {
                                atomicAdd(p.stats + ((size_t)n * p.Cstat + ch) * 2, (double)fs);
                                atomicAdd(p.stats + ((size_t)n * p.Cstat + ch) * 2 + 1, (double)fq);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        DBG_ACC(9, t_all);
        if (p.debug && etid == 0) for (int i = 8; i < 10; i++) atomicAdd(&g_conv_dbg[i], dbg[i]);
        if (p.stats) {
            named_bar_sync(1, kEpiThreads);
            if (cur_n >= 0) flush_stats(p, cta_stats, cur_n, cur_nt, etid);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// L2 promotion of the activation tensor maps.  A halo-tile row is 10 voxels = 160 bytes starting 16 bytes before a
// 128-byte boundary: with 256-byte promotion every such row pulls TWO 256-byte blocks from DRAM (ncu on the dominant layer:
// 422 MB read for 134 MB of input, profiles/r02_ncu_zs_concat_*.csv).  E3B_TMA_PROMO = 0 none, 1 64 B (default: the dominant conv
// launch 120.7 -> 117.7 us in its training form, 104.2 -> 100.2 us in its inference form against 128 B on the same box,
// profiles/r02_tma_promo_ab.txt), 2 128 B, 3 256 B.
CUtensorMapL2promotion tma_l2_promotion()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("E3B_TMA_PROMO"); v = e ? atoi(e) : 1; if (v < 0 || v > 3) v = 1; }
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
         : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}

// QP tensor (N, Cq, Da, Ha, Wa, 4) viewed as 5D (W*4, H, D, Cq, N); box = (bx*4, by, bz, bcq, 1).
// (D, H, W) are the extents of the VIEW starting at `ptr` (a centre-cropped skip tensor is a sub-box of
// its allocation: out-of-view voxels read as 0, exactly like zero padding of the cropped tensor);
// (Da, Ha, Wa) are the extents of the allocation and give the strides.
int make_qp_tensor_map(CUtensorMap* map, const float* ptr, int N, int Cq, int D, int H, int W, int Da, int Ha, int Wa,
                       int bx, int by, int bz, int bcq)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) return set_error("cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[5] = {(cuuint64_t)W * 4, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)Cq, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Wa * 16, (cuuint64_t)Wa * Ha * 16, (cuuint64_t)Wa * Ha * Da * 16,
                             (cuuint64_t)Wa * Ha * Da * 16 * Cq};
    cuuint32_t box[5] = {(cuuint32_t)bx * 4, (cuuint32_t)by, (cuuint32_t)bz, (cuuint32_t)bcq, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (box[0] > 256 || box[1] > 256 || box[2] > 256) return set_error("TMA box dimension > 256");
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, tma_l2_promotion(),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

int conv_ntile_width(int npad_total)
{
    if (npad_total <= 256) return npad_total;
    if (npad_total % 256 == 0) return 256;
    if (npad_total % 128 == 0) return 128;
    if (npad_total % 64 == 0) return 64;
    return -1;
}

// Per-device state: kernel attributes (cudaFuncSetAttribute) and the SM count belong to a DEVICE, not to the
// process: a second device used from the same process (nn.DataParallel replicas, model.to('cuda:1')) needs its own.
int current_device()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

int num_sms()
{
    static int sms[kMaxDevices] = {0};
    const int dev = current_device();
    if (!sms[dev]) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = v > 0 ? v : 148;
    }
    return sms[dev];
}

int conv_debug_read(unsigned long long* out16, int reset)
{
    cudaError_t e = cudaMemcpyFromSymbol(out16, g_conv_dbg, sizeof(unsigned long long) * 16);
    if (e != cudaSuccess) return set_error("conv_debug_read: %s", cudaGetErrorString(e));
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_conv_dbg, z, sizeof(z));
    }
    return 0;
}

// dynamic shared memory the kernel may request: the 227 KB opt-in limit minus its static shared memory
static int conv_max_dyn_smem()
{
    static int v = 0;
    if (!v) {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, conv_tc_kernel) != cudaSuccess) return 0;
        v = 227 * 1024 - (int)fa.sharedSizeBytes;
    }
    return v;
}

int launch_conv_tc(const e3b_conv_args* a, cudaStream_t stream)
{
    ConvTcParams p;
    memset(&p, 0, sizeof(p));
    const int ntaps = a->kd * a->kh * a->kw;
    if (!((a->kd == 1 || a->kd == 3) && (a->kh == 1 || a->kh == 3) && (a->kw == 1 || a->kw == 3)))
        return set_error("conv: taps per dim must be 1 or 3");
    p.N = a->N;
    p.Do = a->D + 2 * a->pd - a->kd + 1;
    p.Ho = a->H + 2 * a->ph - a->kh + 1;
    p.Wo = a->W + 2 * a->pw - a->kw + 1;
    if (p.Do <= 0 || p.Ho <= 0 || p.Wo <= 0) return set_error("conv: empty output");
    p.kd = a->kd; p.kh = a->kh; p.kw = a->kw; p.pd = a->pd; p.ph = a->ph; p.pw = a->pw;
    const int npad_total = a->n_total;          // padded N space (multiple of 16)
    if (npad_total % 16) return set_error("conv: n_total must be a multiple of 16");
    p.NT = conv_ntile_width(npad_total);
    if (p.NT <= 0) return set_error("conv: unsupported output width %d", npad_total);
    p.n_ntiles = npad_total / p.NT;
    // accumulator planes per tile: two ping-pong sets of <= 256 TMEM columns
    int tz = 256 / p.NT; if (tz > 8) tz = 8; if (tz > p.Do) tz = p.Do; if (tz < 1) tz = 1;
    // tile depth: minimise (waves of the persistent schedule) x (cost of a tile ~ tz planes + fixed part)
    {
        const long long txy = (long long)((p.Wo + kTX - 1) / kTX) * ((p.Ho + kTY - 1) / kTY) * p.n_ntiles * a->N;
        double best = 1e30;
        int best_tz = tz;
        for (int t = tz; t >= 1; t--) {
            const long long tiles = txy * ((p.Do + t - 1) / t);
            const long long waves = (tiles + num_sms() - 1) / num_sms();
            const double cost = (double)waves * (t + 0.5);
            if (cost < best - 1e-9) { best = cost; best_tz = t; }
        }
        tz = best_tz;
    }
    if (a->force_tz > 0) tz = a->force_tz;
    p.TZ = tz;
    p.HX = kTX + a->kw - 1; p.HY = kTY + a->kh - 1; p.HZ = tz + a->kd - 1;
    p.tiles_x = (p.Wo + kTX - 1) / kTX; p.tiles_y = (p.Ho + kTY - 1) / kTY; p.tiles_z = (p.Do + tz - 1) / tz;
    p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_z * p.n_ntiles * a->N;
    p.chunks0 = e3b_cpad16(a->C0) / 16;
    p.chunks1 = a->src1 ? e3b_cpad16(a->C1) / 16 : 0;
    p.off1_d = a->off1_d; p.off1_h = a->off1_h; p.off1_w = a->off1_w;
    p.a_stage_bytes = (uint32_t)(p.HX * p.HY * p.HZ * 16 * 2);
    // weight stage: the largest tap group (27, 9, 3, 1) that still leaves room for two input and two weight stages.
    // Every stage boundary costs the issuing thread an mbarrier wait (~116 cycles) and a tcgen05.commit (~190 cycles of
    // pipe time; profiles/r02_umma_issue_queue_commit_wait.txt), so fewer, larger stages win as long as both rings stay
    // double-buffered.
    if (conv_max_dyn_smem() <= 0) return set_error("conv: cudaFuncGetAttributes failed");
    const size_t budget = (size_t)conv_max_dyn_smem() - 1024 - 256;     // barriers + alignment slack
    int tg = ntaps;
    while (tg > 1 && 2 * (size_t)p.a_stage_bytes + 2 * ((size_t)tg * 2 * p.NT * 16) > budget) tg /= 3;
    if (const char* e = getenv("E3B_CONV_TG")) { const int v = atoi(e); if (v >= 1 && v <= tg && ntaps % v == 0) tg = v; }     // tuning
    p.TG = tg;
    p.b_stage_bytes = (uint32_t)(tg * 2 * p.NT * 16);
    int sa = 2, sb = 2;
    // grow depth while it fits (A first up to 4, then B up to 4)
    for (;;) {
        bool grew = false;
        if (sa < 4 && (size_t)(sa + 1) * p.a_stage_bytes + (size_t)sb * p.b_stage_bytes <= budget) { sa++; grew = true; }
        if (sb < 4 && (size_t)sa * p.a_stage_bytes + (size_t)(sb + 1) * p.b_stage_bytes <= budget) { sb++; grew = true; }
        if (!grew) break;
    }
    if ((size_t)sa * p.a_stage_bytes + (size_t)sb * p.b_stage_bytes > budget)
        return set_error("conv: tile does not fit shared memory");
    p.SA = sa; p.SB = sb;
    p.bias = a->bias; p.n_bias = a->n_bias;
    p.dst0 = reinterpret_cast<float*>(a->dst0); p.dst1 = a->dst1;
    p.cq0_alloc = a->half_out ? e3b_cpad16(a->Cd0) / 8 : e3b_cpad(a->Cd0) / 4;      // 16-byte planes in dst0
    p.cq1_alloc = a->dst1 ? e3b_cpad(a->Cd1) / 4 : 0;
    p.cq0 = a->dst1 ? p.cq0_alloc : (1 << 30);
    p.relu = a->relu; p.half_out = a->half_out; p.out_scale = a->out_scale; p.w_unscale = a->w_unscale;
    if (a->half_out && a->dst1) return set_error("conv: the fp16 operand output has a single destination");
    if (a->half_out && a->stats) return set_error("conv: statistics are taken from an fp32 output");
    p.debug = getenv("E3B_CONV_DEBUG") != nullptr;
    p.stats = a->stats; p.Cstat = a->stats_channels;
    p.scatter = a->scatter; p.sd = a->sd; p.sh = a->sh; p.sw = a->sw;
    p.Cup = a->scatter ? e3b_cpad16(a->Cd0) : 1;
    p.Ds = a->Ds; p.Hs = a->Hs; p.Ws = a->Ws;
    p.wpk = reinterpret_cast<const float*>(a->wpk);

    CUtensorMap m0, m1;
    int rc = make_qp_tensor_map(&m0, reinterpret_cast<const float*>(a->src0), a->N, p.chunks0 * 2, a->D, a->H, a->W, a->D, a->H, a->W, p.HX, p.HY,
                                p.HZ, 2);
    if (rc) return rc;
    if (a->src1) {
        const float* v1 = reinterpret_cast<const float*>(a->src1) + (((size_t)a->off1_d * a->H1 + a->off1_h) * a->W1 + a->off1_w) * 4;
        rc = make_qp_tensor_map(&m1, v1, a->N, p.chunks1 * 2, a->D, a->H, a->W, a->D1, a->H1, a->W1, p.HX, p.HY, p.HZ, 2);
        if (rc) return rc;
    } else {
        m1 = m0;
    }
    const size_t smem = (size_t)sa * p.a_stage_bytes + (size_t)sb * p.b_stage_bytes + 1024;
    static bool configured[kMaxDevices] = {false};
    if (!configured[current_device()]) {
        // static shared memory (cta_stats) counts against the 227 KB per-CTA limit
        const int max_dyn = conv_max_dyn_smem();
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
        if (e != cudaSuccess) return set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured[current_device()] = true;
    }
    if (p.stats) {
        cudaError_t e = cudaMemsetAsync(p.stats, 0, sizeof(double) * 2 * (size_t)a->N * p.Cstat, stream);
        if (e != cudaSuccess) return set_error("conv: stats memset: %s", cudaGetErrorString(e));
    }
    int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    conv_tc_kernel<<<grid, kThreads, smem, stream>>>(m0, m1, p);
    return check_launch("conv_tc");
}

}  // namespace e3b
