// HBM-bound kernels of the UNet path: layout conversion, weight packing, norm statistics finalisation,
// fused normalise + ReLU (+ max-pool), their backward (with un-pooling, skip-gradient add and
// space-to-depth output), and the 1x1x1 head fused with softmax / argmax / crop-and-place.
// All of them move 16-byte quads with consecutive threads on consecutive voxels (coalesced 512 B/warp).
#include "common.cuh"
#include "kernels.h"

namespace e3b {

static inline int grid_for(size_t total, int block, int max_waves = 8)
{
    size_t b = (total + block - 1) / block;
    size_t cap = (size_t)num_sms() * max_waves * (2048 / block);
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// fp16 operand tensor (QH): (N, Ch, S voxels, 8 halves), Ch = ceil16(C)/8.  The fp32 quad `cq` (channels
// 4cq..4cq+3) of voxel v is the 8-byte half (cq & 1) of unit ((n*Ch + cq/2)*S + v).
E3B_DEVINL size_t qh_index(int n, int Ch, int cq, size_t S, size_t v) { return (((size_t)n * Ch + (cq >> 1)) * S + v) * 2 + (cq & 1); }
E3B_DEVINL void store_qh(uint2* __restrict__ qh, const float4& v, int n, int Ch, int cq, size_t S, size_t vox) {
    qh[qh_index(n, Ch, cq, S, vox)] = pack_half4(v.x, v.y, v.z, v.w);
}

// ------------------------------------------------------------------------------------------------
// NCDHW box -> QH   (network input, Predictor tile gather)
// ------------------------------------------------------------------------------------------------
// grid: (voxel chunks, 8-channel planes, N); one thread writes one 16-byte unit (8 channels of one voxel): fully
// coalesced stores, 32-bit index arithmetic (the first version decoded a flat 64-bit index per 8-byte half: 44 us for the
// 4 MB input of cfg 2, now memory bound).  Padding channels of the 16-channel chunks are written as 0.
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ src, const int32_t* __restrict__ origins,
                                                   uint4* __restrict__ dst, int C, int Ch, int D, int H, int W, int Dv, int Hv,
                                                   int Wv, int z0, int y0, int x0, int single, int flip)
{
    const int S = D * H * W, HW = H * W;
    const int ch = blockIdx.y, n = blockIdx.z;
    int oz = z0, oy = y0, ox = x0;
    if (origins) { oz = origins[3 * n]; oy = origins[3 * n + 1]; ox = origins[3 * n + 2]; }
    const size_t plane = (size_t)Dv * Hv * Wv;
    const float* sb = src + (single ? 0 : (size_t)n * C * plane) + (size_t)ch * 8 * plane;
    const int nc = min(8, C - ch * 8);                  // real channels in this unit (<= 0: all padding)
    if (nc <= 0) {                                      // a plane of padding channels only: zero fill, no index arithmetic
        for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < S; v += gridDim.x * blockDim.x)
            dst[((size_t)n * Ch + ch) * S + v] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < S; v += gridDim.x * blockDim.x) {
        const int z = v / HW, r = v - z * HW, y = r / W, x = r - y * W;
        // (flip: the tile is mirrored while it is gathered -- FlipAugment.forward of the Predictor's TTA)
        const int sz = ((flip & 1) ? D - 1 - z : z) + oz, sy = ((flip & 2) ? H - 1 - y : y) + oy,
                  sx = ((flip & 4) ? W - 1 - x : x) + ox;
        float q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (nc > 0 && sz >= 0 && sz < Dv && sy >= 0 && sy < Hv && sx >= 0 && sx < Wv) {
            const size_t vox = ((size_t)sz * Hv + sy) * Wv + sx;
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < nc) q[j] = __ldg(sb + (size_t)j * plane + vox);
        }
        const uint2 lo = pack_half4(tf32_rn(q[0]), tf32_rn(q[1]), tf32_rn(q[2]), tf32_rn(q[3]));
        const uint2 hi = pack_half4(tf32_rn(q[4]), tf32_rn(q[5]), tf32_rn(q[6]), tf32_rn(q[7]));
        dst[((size_t)n * Ch + ch) * S + v] = make_uint4(lo.x, lo.y, hi.x, hi.y);
    }
}

__global__ void unpack_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int C, int Cq, size_t S)
{
    const size_t total = (size_t)N * C * S;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t v = i % S;
        const int c = (int)((i / S) % C);
        const int n = (int)(i / (S * C));
        dst[i] = src[(((size_t)n * Cq + (c >> 2)) * S + v) * 4 + (c & 3)];
    }
}

// ------------------------------------------------------------------------------------------------
// weight packing: torch layout -> fp16 [ntile][chunk16][tap][kq 2][NT][8] (K-major no-swizzle smem image)
// ------------------------------------------------------------------------------------------------
struct PackDims { int ktot, ntot, taps, NT; };

static PackDims pack_dims(int mode, int C0, int C1, int Co, int kd, int kh, int kw)
{
    PackDims d;
    const int t = kd * kh * kw;
    switch (mode) {
    // z-stacked images: one image per N tile (conv_zs_ntile: the whole width, or tiles of 32 columns)
    case 4: d.ktot = cpad16(C0) + (C1 > 0 ? cpad16(C1) : 0); d.ntot = cpad16(Co); d.taps = t; d.NT = conv_zs_ntile(C0, C1, d.ntot); return d;
    case 5: d.ktot = cpad16(Co); d.ntot = cpad16(cpad8(C0) + (C1 > 0 ? cpad8(C1) : 0)); d.taps = t; d.NT = conv_zs_ntile(Co, 0, d.ntot); return d;
    case 0: d.ktot = cpad16(C0) + (C1 > 0 ? cpad16(C1) : 0); d.ntot = cpad16(Co); d.taps = t; break;
    case 1: d.ktot = cpad16(Co); d.ntot = cpad16(cpad8(C0) + (C1 > 0 ? cpad8(C1) : 0)); d.taps = t; break;
    case 2: d.ktot = cpad16(C0); d.ntot = t * cpad16(Co); d.taps = 1; break;
    default: d.ktot = cpad16(t * cpad8(Co)); d.ntot = cpad16(C0); d.taps = 1; break;
    }
    d.NT = conv_ntile_width(d.ntot);
    return d;
}

// one element of a packed weight image (shared by the single-image and the batched kernel)
E3B_DEVINL __half pack_weight_element(int mode, const float* __restrict__ w, const float* __restrict__ scale, float ws, int C0, int C1,
                                      int Co, int tu, const PackDims& d, size_t i)
{
    const int C0p16 = cpad16(C0), C0p8 = cpad8(C0), Cop8 = cpad8(Co), Cop16 = cpad16(Co);
    const int nchunks = d.ktot / 16;
    size_t r = i;
    const int k8 = (int)(r % 8); r /= 8;
    int nn, kq, tap, chunk, nt = 0;
    if (mode >= 4) {
        // z-stacked image (conv_zs.cu): [N tile][chunk16][tap (dy,dx) 9][kq 2][N3 = (j, n)][8], block j holds z tap dz = 2 - j
        const int n3 = (int)(r % (3 * d.NT)); r /= (size_t)(3 * d.NT);
        kq = (int)(r % 2); r /= 2;
        const int t9 = (int)(r % 9); r /= 9;
        chunk = (int)(r % nchunks);
        nt = (int)(r / nchunks);
        nn = n3 % d.NT;
        tap = (2 - n3 / d.NT) * 9 + t9;
    } else {
        nn = (int)(r % d.NT); r /= d.NT;
        kq = (int)(r % 2); r /= 2;
        tap = (int)(r % d.taps); r /= d.taps;
        chunk = (int)(r % nchunks);
        nt = (int)(r / nchunks);
    }
    const int k = chunk * 16 + kq * 8 + k8;
    const int n = nt * d.NT + nn;
    float v = 0.f;
    if (mode == 0 || mode == 4) {
        // K space [pad16(C0) | pad16(C1)] (the operand tensors of the two sources), N = output channel
        int ci = -1;
        if (k < C0p16) { if (k < C0) ci = k; }
        else if (k - C0p16 < C1) ci = C0 + k - C0p16;
        if (ci >= 0 && n < Co) {
            v = w[((size_t)n * (C0 + C1) + ci) * d.taps + tap];
            if (scale) v *= scale[n];
        }
    } else if (mode == 1 || mode == 5) {
        // K = output channel of the forward conv, N space [pad8(C0) | pad8(C1)] (the fp32 QP gradient outputs)
        int ci = -1;
        if (n < C0p8) { if (n < C0) ci = n; }
        else if (n - C0p8 < C1) ci = C0 + n - C0p8;
        if (ci >= 0 && k < Co) v = w[((size_t)k * (C0 + C1) + ci) * d.taps + (d.taps - 1 - tap)];
    } else if (mode == 2) {
        const int t = n / Cop16, co = n % Cop16;
        if (k < C0 && co < Co) v = w[((size_t)k * Co + co) * tu + t];
    } else {
        const int t = k / Cop8, co = k % Cop8;
        if (t < tu && n < C0 && co < Co) v = w[((size_t)n * Co + co) * tu + t];
    }
    return __float2half_rn(v * ws);
}

__global__ void pack_weights_kernel(int mode, const float* __restrict__ w, const float* __restrict__ scale,
                                    const float* __restrict__ wscale, __half* __restrict__ dst, int C0, int C1, int Co, int tu,
                                    PackDims d)
{
    const float ws = wscale ? __ldg(wscale) : 1.f;       // power of two: max|w| -> [1, 2), undone in the conv epilogue
    const size_t total = (size_t)d.ktot * d.ntot * d.taps;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = pack_weight_element(mode, w, scale, ws, C0, C1, Co, tu, d, i);
}

// All weight images of a network in ONE launch (a training step re-packs every image: 23 launches for BASELINE cfg 2):
// blocks of 256 elements, each block finds its job in the (block-prefix-summed) job table.
struct PackJobDev {
    const float* w; const float* scale; const float* wscale; __half* dst;
    int mode, C0, C1, Co, tu, first_block;
    PackDims d;
    unsigned long long total;
};

__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const PackJobDev* __restrict__ jobs, int njobs)
{
    int lo = 0, hi = njobs - 1;                           // last job whose first_block <= blockIdx.x
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].first_block <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const PackJobDev j = jobs[lo];
    const size_t i = (size_t)(blockIdx.x - j.first_block) * 256 + threadIdx.x;
    if (i >= j.total) return;
    const float ws = j.wscale ? __ldg(j.wscale) : 1.f;
    j.dst[i] = pack_weight_element(j.mode, j.w, j.scale, ws, j.C0, j.C1, j.Co, j.tu, j.d, i);
}

// per-tensor power-of-two weight scales: grid (chunks, tensors); table[b] = (2^k, 2^-k) with k = -floor(log2(max |w|))
// (max |w| lands in [1, 2); 0 / non-finite maxima give k = 0), |k| <= 100.  scratch[2b] collects the maximum (float bits of
// a non-negative float order like unsigned integers), scratch[2b + 1] counts finished blocks; the last block of a tensor
// writes the table row and leaves both words zero for the next call.
static constexpr int kWsChunks = 16;
__global__ void __launch_bounds__(256) weight_scales_kernel(const e3b_ws_job* __restrict__ jobs, float* __restrict__ table,
                                                            unsigned int* __restrict__ scratch)
{
    const int b = blockIdx.y;
    const e3b_ws_job j = jobs[b];
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < j.n; i += 256ll * kWsChunks) {
        const float v = fabsf(__ldg(j.w + i));
        m = (v <= 3.4e38f) ? fmaxf(m, v) : INFINITY;            // NaN / inf poison the maximum
    }
    __shared__ float red[8];
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 0; w < 8; w++) m = fmaxf(m, red[w]);
        atomicMax(&scratch[2 * b], __float_as_uint(m));
        __threadfence();
        if (atomicAdd(&scratch[2 * b + 1], 1u) == (unsigned)gridDim.x - 1) {
            __threadfence();
            const float amax = __uint_as_float(atomicExch(&scratch[2 * b], 0u));
            scratch[2 * b + 1] = 0u;
            float k = 0.f;
            if (amax > 0.f && amax <= 3.4e38f) k = -floorf(log2f(fmaxf(amax, 1e-37f)));
            k = fminf(fmaxf(k, -100.f), 100.f);
            table[2 * b] = exp2f(k);
            table[2 * b + 1] = exp2f(-k);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// normalisation statistics -> per-(n,c) scale / shift (+ running stats)
// ------------------------------------------------------------------------------------------------
__global__ void norm_finalize_kernel(const double* __restrict__ stats, int mode, int G, int N, int C, int Cp, double S,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                     float* running_mean, float* running_var, float momentum, float* __restrict__ scale,
                                     float* __restrict__ shift, float* __restrict__ mean_o, float* __restrict__ rstd_o)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * Cp) return;
    const int n = i / Cp, c = i % Cp;
    float sc = 0.f, sh = 0.f, mu_f = 0.f, r_f = 1.f;
    if (c < C) {
        const double ga = gamma ? (double)gamma[c] : 1.0, be = beta ? (double)beta[c] : 0.0;
        double mu = 0.0, var = 1.0 - (double)eps;
        if (mode == 0) {
            sc = 1.f; sh = 0.f;
        } else {
            if (mode == 1) {
                const int cg = C / G, g = c / cg;
                double s = 0.0, ss = 0.0;
                for (int j = 0; j < cg; j++) {
                    s += stats[((size_t)n * C + g * cg + j) * 2];
                    ss += stats[((size_t)n * C + g * cg + j) * 2 + 1];
                }
                const double cnt = S * cg;
                mu = s / cnt; var = ss / cnt - mu * mu;
            } else if (mode == 2) {
                double s = 0.0, ss = 0.0;
                for (int j = 0; j < N; j++) { s += stats[((size_t)j * C + c) * 2]; ss += stats[((size_t)j * C + c) * 2 + 1]; }
                const double cnt = S * N;
                mu = s / cnt; var = ss / cnt - mu * mu;
                if (n == 0 && running_mean) {
                    const double unb = cnt > 1 ? var * cnt / (cnt - 1.0) : var;
                    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mu);
                    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
                }
            } else {
                mu = running_mean[c]; var = running_var[c];
            }
            if (var < 0.0) var = 0.0;
            const double r = 1.0 / sqrt(var + (double)eps);
            sc = (float)(r * ga); sh = (float)(be - mu * r * ga);
            mu_f = (float)mu; r_f = (float)r;
        }
    }
    scale[i] = sc; shift[i] = sh;
    if (mean_o) mean_o[i] = mu_f;
    if (rstd_o) rstd_o[i] = r_f;
}

// ------------------------------------------------------------------------------------------------
// a = relu(y*scale+shift)  [+ pooled = maxpool(a), kernel (pkd,pkh,pkw), ceil mode]
// one thread per pooling window (or voxel) per 4-channel plane
// ------------------------------------------------------------------------------------------------
// grid: (planes * chunks-per-plane, Cq, N): one block works inside one (n, 4-channel plane, z) slice, so
// the per-(n,c) constants are loaded once and only one integer division per thread is needed.
struct NormActDev {
    const float4* y; const float* scale; const float* shift;
    const uint2* yh;             // y_half: the input is itself a QH operand tensor (eval path: pooling only)
    uint2* a; uint2* pooled;     // QH outputs
    uchar4* pool_idx;
    int C, N, Cq, Ch, D, H, W, pkd, pkh, pkw, relu;
    float slope;                 // activation code / negative slope, see act_fwd
    const float* slope_dev;      // nn.PReLU: the (learned) negative slope lives in device memory
    int Dp, Hp, Wp;
};

// Activation codes (e3b_norm_act `relu`, e3b_norm_bwd_args.relu): 0 identity ('lin'), 1 the leaky-ReLU family with the
// negative slope `act_slope` (0 = ReLU, 0.1 = 'leaky', RReLU's eval slope ...), 2 SiLU  (get_activation, unet.py:183-199).
E3B_DEVINL float act_fwd(float z, int act, float slope)
{
    if (act == 1) return slope == 0.f ? fmaxf(z, 0.f) : (z > 0.f ? z : z * slope);
    if (act == 2) return z / (1.f + __expf(-z));
    return z;
}
// upstream gradient g times the activation's derivative at the pre-activation z
E3B_DEVINL float act_bwd(float z, float g, int act, float slope)
{
    if (act == 1) return z > 0.f ? g : (slope == 0.f ? 0.f : g * slope);
    if (act == 2) { const float s = 1.f / (1.f + __expf(-z)); return g * s * fmaf(z, 1.f - s, 1.f); }
    return g;
}

E3B_DEVINL float4 norm_affine(const float4& yv, const float4& sc, const float4& sh, bool affine)
{
    float4 v = yv;
    if (affine) {
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
        v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    }
    return v;
}

E3B_DEVINL float4 norm_relu_round(const float4& yv, const float4& sc, const float4& sh, bool affine, int act, float slope)
{
    float4 v = norm_affine(yv, sc, sh, affine);
    v.x = act_fwd(v.x, act, slope); v.y = act_fwd(v.y, act, slope); v.z = act_fwd(v.z, act, slope); v.w = act_fwd(v.w, act, slope);
    // activations are MMA operands of the next conv: store them rounded to TF32
    v.x = tf32_rn(v.x); v.y = tf32_rn(v.y); v.z = tf32_rn(v.z); v.w = tf32_rn(v.w);
    return v;
}

E3B_DEVINL void load_nc4(const float* p, size_t nc, float4& v, float dflt) {
    v = p ? *reinterpret_cast<const float4*>(p + nc) : make_float4(dflt, dflt, dflt, dflt);
}

// a = tf32(relu(y*scale+shift)), no pooling.  Each thread owns kVpt voxels (a block covers kVpt*256
// consecutive voxels of one (n, 4-channel plane) slab) and issues all its loads before the first use:
// ~64 B in flight per thread is what it takes to cover the HBM latency at 50 % occupancy.
static constexpr int kVpt = 4;
__global__ void __launch_bounds__(256) norm_act_kernel(const NormActDev p)
{
    const int cq = blockIdx.y, n = blockIdx.z;
    const int S = p.D * p.H * p.W;
    const int v0 = blockIdx.x * (256 * kVpt) + threadIdx.x;
    const size_t nc = ((size_t)n * p.Cq + cq) * 4;
    float4 sc, sh;
    load_nc4(p.scale, nc, sc, 1.f); load_nc4(p.shift, nc, sh, 0.f);
    const size_t base = ((size_t)n * p.Cq + cq) * (size_t)S;
    const float slope = p.slope_dev ? __ldg(p.slope_dev) : p.slope;
    float4 yv[kVpt];
#pragma unroll
    for (int j = 0; j < kVpt; j++) {
        const int v = v0 + j * 256;
        if (v < S) yv[j] = __ldcs(p.y + base + v);
    }
#pragma unroll
    for (int j = 0; j < kVpt; j++) {
        const int v = v0 + j * 256;
        if (v >= S) continue;
        const float4 r = norm_relu_round(yv[j], sc, sh, p.scale != nullptr, p.relu, slope);
        if (p.a) store_qh(p.a, r, n, p.Ch, cq, (size_t)S, (size_t)v);
    }
}

// ... + ceil-mode max pooling: one pooling window per thread.  pool_idx records, per channel, the window
// slot ((dz*pkh + dy)*pkw + dx) of the FIRST maximum in scan order (torch max_pool backward semantics), so
// that the backward pass is a per-voxel kernel that never re-examines the window.
__global__ void __launch_bounds__(256) norm_act_pool_kernel(const NormActDev p)
{
    const int cq = blockIdx.y, n = blockIdx.z;
    const int chunks = (p.Hp * p.Wp + 255) / 256;
    const int zp = blockIdx.x / chunks;
    const int hw = (blockIdx.x % chunks) * 256 + threadIdx.x;
    if (hw >= p.Hp * p.Wp) return;
    const int yp = hw / p.Wp, xp = hw % p.Wp;
    const size_t nc = ((size_t)n * p.Cq + cq) * 4;
    float4 sc, sh;
    load_nc4(p.scale, nc, sc, 1.f); load_nc4(p.shift, nc, sh, 0.f);
    const size_t base = ((size_t)n * p.Cq + cq) * p.D;
    const size_t S = (size_t)p.D * p.H * p.W;
    const float slope = p.slope_dev ? __ldg(p.slope_dev) : p.slope;
    float4 m = make_float4(-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f);
    uchar4 idx = make_uchar4(0, 0, 0, 0);
#pragma unroll
    for (int dz = 0; dz < 2; dz++) {
        const int z = zp * p.pkd + dz;
        if (dz >= p.pkd || z >= p.D) continue;
#pragma unroll
        for (int dy = 0; dy < 2; dy++) {
            const int yy = yp * p.pkh + dy;
            if (dy >= p.pkh || yy >= p.H) continue;
#pragma unroll
            for (int dx = 0; dx < 2; dx++) {
                const int x = xp * p.pkw + dx;
                if (dx >= p.pkw || x >= p.W) continue;
                const size_t vox = ((size_t)z * p.H + yy) * p.W + x;
                float4 v;
                if (p.yh) v = unpack_half4(p.yh[qh_index(n, p.Ch, cq, S, vox)]);
                else v = norm_relu_round(p.y[base * p.H * p.W + vox], sc, sh, p.scale != nullptr, p.relu, slope);
                if (p.a) store_qh(p.a, v, n, p.Ch, cq, S, vox);
                const unsigned char slot = (unsigned char)((dz * p.pkh + dy) * p.pkw + dx);
                if (v.x > m.x) { m.x = v.x; idx.x = slot; }
                if (v.y > m.y) { m.y = v.y; idx.y = slot; }
                if (v.z > m.z) { m.z = v.z; idx.z = slot; }
                if (v.w > m.w) { m.w = v.w; idx.w = slot; }
            }
        }
    }
    const size_t op = ((((size_t)n * p.Cq + cq) * p.Dp + zp) * p.Hp + yp) * p.Wp + xp;
    if (p.pooled) store_qh(p.pooled, m, n, p.Ch, cq, (size_t)p.Dp * p.Hp * p.Wp, ((size_t)zp * p.Hp + yp) * p.Wp + xp);
    if (p.pool_idx) p.pool_idx[op] = idx;
}

// Pooling alone of a QH activation (inference: the conv epilogue already wrote the activated fp16 tensor, nothing is kept
// for a backward pass): one thread per pooled voxel and 8-channel unit, whole 16-byte units in and out (the generic kernel
// above reads 8-byte halves at a 32-byte stride: 3.3 TB/s on the 1 GB level-0 tensor of a 32-tile Predictor pass).
// grid: (chunks of pooled voxels, Ch, N)
__global__ void __launch_bounds__(256) pool_qh_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int Ch, int D, int H, int W,
                                                      int pkd, int pkh, int pkw, int Dp, int Hp, int Wp)
{
    const int ch = blockIdx.y, n = blockIdx.z;
    const int Sp = Dp * Hp * Wp, HWp = Hp * Wp;
    const uint4* sb = src + ((size_t)n * Ch + ch) * ((size_t)D * H * W);
    uint4* db = dst + ((size_t)n * Ch + ch) * (size_t)Sp;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < Sp; v += gridDim.x * blockDim.x) {
        const int zp = v / HWp, r = v - zp * HWp, yp = r / Wp, xp = r - yp * Wp;
        __half2 m[4];
        bool first = true;
        for (int dz = 0; dz < pkd; dz++) {
            const int z = zp * pkd + dz;
            if (z >= D) break;
            for (int dy = 0; dy < pkh; dy++) {
                const int y = yp * pkh + dy;
                if (y >= H) break;
                for (int dx = 0; dx < pkw; dx++) {
                    const int x = xp * pkw + dx;
                    if (x >= W) break;
                    const uint4 u = __ldg(sb + ((size_t)z * H + y) * W + x);
                    const __half2* h = reinterpret_cast<const __half2*>(&u);
                    if (first) { m[0] = h[0]; m[1] = h[1]; m[2] = h[2]; m[3] = h[3]; first = false; }
                    else { m[0] = __hmax2(m[0], h[0]); m[1] = __hmax2(m[1], h[1]); m[2] = __hmax2(m[2], h[2]); m[3] = __hmax2(m[3], h[3]); }
                }
            }
        }
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
        oh[0] = m[0]; oh[1] = m[1]; oh[2] = m[2]; oh[3] = m[3];
        db[v] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// backward of norm -> relu [-> pool]: per-voxel kernels
// ------------------------------------------------------------------------------------------------
struct NormBwdDev {
    const float4 *y, *g0, *g1, *gp;
    const uchar4* pool_idx;
    const float *scale, *shift;              // forward affine (nullptr: y already is the activation)
    int N, Cq, C, D, H, W, wd, wh, ww;       // window = pooling kernel (gp) or s2d stride, else 1
    int Dw, Hw, Ww;
    int Dg, Hg, Wg;                          // extents of the thread grid (s2d: the un-cropped fine grid)
    int relu, s2d;
    float slope;                             // activation code / negative slope, see act_fwd
    const float* slope_dev;                  // nn.PReLU: the slope in device memory (three-kernel path only)
    double* slope_sums;                      // [N][pad8(C)] sum over z <= 0 of g * z: the slope's gradient (reduce pass)
    const float *gamma, *mean, *rstd, *m1, *m2;
    double* sums;
    unsigned int* amax;                      // [N][pad8(C)][2] max |dr|, max |xhat| (float bits; reduce pass)
    float* dy_scale;                         // [0] bound on |dy| (float bits), [1] 2^k, [2] 2^-k
    uint2* dy;                               // QH output, scaled by 2^k
    int Ch;                                  // 16-byte planes of dy
    int g1_crop, g1_od, g1_oh, g1_ow, g1_D, g1_H, g1_W;   // g1 = gradient of a centre-cropped view (zero outside the box)
};

// power-of-two scale that brings a tensor bounded by `bound` to at most 2^14 (fp16 max is 2^16)
E3B_DEVINL float dy_scale_from_bound(float bound) {
    if (!(bound > 0.f) || !(bound < 3.0e38f)) return 1.f;
    int e = 14 - (int)ceilf(log2f(bound));
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    return exp2f((float)e);
}

// Inputs of one voxel of the backward pass, loaded up front (several voxels per thread are in flight
// before the first one is used).
struct VoxIn {
    float4 y, g0, g1, gp;
    uchar4 idx;
    unsigned char slot;
};

// STREAM: last use of the gradient tensors (the apply pass): evict-first loads.  The reduce pass loads them with the
// default policy so that the apply pass, which walks the tensor in the opposite order, finds the tail in the 126 MB L2.
template <bool STREAM>
E3B_DEVINL float4 ld_grad(const float4* p) { return STREAM ? __ldcs(p) : __ldg(p); }

// GENERAL = false: the caller knows there is no pooled gradient and no cropped skip gradient
template <bool STREAM, bool GENERAL = true>
E3B_DEVINL void load_vox(const NormBwdDev& p, size_t o, int n, int cq, int z, int yy, int x, VoxIn& in)
{
    // every field the instantiation can read gets a value: a conditionally written field of a loop-local struct would
    // otherwise live in local memory (ptxas keeps the "old" value across the predicated load)
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    in.y = p.y[o];
    in.g0 = zero; in.g1 = zero;
    if (GENERAL) { in.gp = zero; in.idx = make_uchar4(0, 0, 0, 0); in.slot = 255; }
    if (p.g0) in.g0 = ld_grad<STREAM>(p.g0 + o);
    if (p.g1) {
        if (!GENERAL || !p.g1_crop) {
            in.g1 = ld_grad<STREAM>(p.g1 + o);
        } else {
            // backward of autocrop's slice of the skip tensor: zero outside the cropped box
            const int zc = z - p.g1_od, yc = yy - p.g1_oh, xc = x - p.g1_ow;
            in.g1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (zc >= 0 && zc < p.g1_D && yc >= 0 && yc < p.g1_H && xc >= 0 && xc < p.g1_W)
                in.g1 = ld_grad<STREAM>(p.g1 + ((((size_t)n * p.Cq + cq) * p.g1_D + zc) * p.g1_H + yc) * p.g1_W + xc);
        }
    }
    if (GENERAL && p.gp) {
        const int zw = z / p.wd, yw = yy / p.wh, xw = x / p.ww;
        in.slot = (unsigned char)(((z - zw * p.wd) * p.wh + (yy - yw * p.wh)) * p.ww + (x - xw * p.ww));
        const size_t ow = ((((size_t)n * p.Cq + cq) * p.Dw + zw) * p.Hw + yw) * p.Ww + xw;
        in.idx = p.pool_idx[ow];
        in.gp = p.gp[ow];
    }
}

// the upstream gradient through the activation, dr = (g0 + g1 + unpool(gp)) * act'(z) ([z > 0] for ReLU), and xhat of one voxel.
// z is recomputed bit-exactly from y (the arithmetic of norm_act_kernel), never read.
template <bool GENERAL = true>
E3B_DEVINL void voxel_grad(const NormBwdDev& p, const VoxIn& in, const float4& mu, const float4& rs, const float4& sc,
                           const float4& sh, float4& dr, float4& xh, float slope, float4* gz = nullptr)
{
    const float4 yv = in.y;
    const float4 zv = norm_affine(yv, sc, sh, p.scale != nullptr);       // the pre-activation, recomputed as the forward did
    xh = make_float4((yv.x - mu.x) * rs.x, (yv.y - mu.y) * rs.y, (yv.z - mu.z) * rs.z, (yv.w - mu.w) * rs.w);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.g0) g = in.g0;
    if (p.g1) { g.x += in.g1.x; g.y += in.g1.y; g.z += in.g1.z; g.w += in.g1.w; }
    if (GENERAL && p.gp) {
        if (in.idx.x == in.slot) g.x += in.gp.x;
        if (in.idx.y == in.slot) g.y += in.gp.y;
        if (in.idx.z == in.slot) g.z += in.gp.z;
        if (in.idx.w == in.slot) g.w += in.gp.w;
    }
    if (gz)      // d a / d slope = z where z <= 0 (nn.PReLU)
        *gz = make_float4(zv.x > 0.f ? 0.f : g.x * zv.x, zv.y > 0.f ? 0.f : g.y * zv.y, zv.z > 0.f ? 0.f : g.z * zv.z,
                          zv.w > 0.f ? 0.f : g.w * zv.w);
    g.x = act_bwd(zv.x, g.x, p.relu, slope); g.y = act_bwd(zv.y, g.y, p.relu, slope);
    g.z = act_bwd(zv.z, g.z, p.relu, slope); g.w = act_bwd(zv.w, g.w, p.relu, slope);
    dr = g;
}

// grid: (chunks, Cq, N); a block strides over the voxels of one (n, 4-channel plane) slab, kRedVpt at a time
static constexpr int kRedVpt = 1;
__global__ void __launch_bounds__(256) norm_bwd_reduce_kernel(const NormBwdDev p)
{
    const int cq = blockIdx.y, n = blockIdx.z;
    const size_t nc = ((size_t)n * p.Cq + cq) * 4;
    float4 mu, rs, sc, sh;
    load_nc4(p.mean, nc, mu, 0.f); load_nc4(p.rstd, nc, rs, 1.f);
    load_nc4(p.scale, nc, sc, 1.f); load_nc4(p.shift, nc, sh, 0.f);
    const int HW = p.H * p.W;
    const int total = p.D * HW;
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    float s3 = 0.f;                                          // nn.PReLU: sum over z <= 0 of g * z (all four channels)
    const float slope = p.slope_dev ? __ldg(p.slope_dev) : p.slope;
    float md[4] = {0, 0, 0, 0}, mx[4] = {0, 0, 0, 0};       // max |dr|, max |xhat|: bound for the fp16 scale of dy
    const size_t base = ((size_t)n * p.Cq + cq) * (size_t)total;
    const int stride = gridDim.x * blockDim.x;
    for (int v0 = blockIdx.x * blockDim.x + threadIdx.x; v0 < total; v0 += kRedVpt * stride) {
        VoxIn in[kRedVpt];
#pragma unroll
        for (int j = 0; j < kRedVpt; j++) {
            const int v = v0 + j * stride;
            if (v < total) {
                const int z = v / HW, r = v - z * HW, yy = r / p.W, x = r - yy * p.W;
                load_vox<false>(p, base + v, n, cq, z, yy, x, in[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < kRedVpt; j++) {
            if (v0 + j * stride >= total) continue;
            float4 dr, xh, gz;
            voxel_grad(p, in[j], mu, rs, sc, sh, dr, xh, slope, p.slope_sums ? &gz : nullptr);
            if (p.slope_sums) s3 += (gz.x + gz.y) + (gz.z + gz.w);
            s1[0] += dr.x; s1[1] += dr.y; s1[2] += dr.z; s1[3] += dr.w;
            s2[0] = fmaf(dr.x, xh.x, s2[0]); s2[1] = fmaf(dr.y, xh.y, s2[1]);
            s2[2] = fmaf(dr.z, xh.z, s2[2]); s2[3] = fmaf(dr.w, xh.w, s2[3]);
            md[0] = fmaxf(md[0], fabsf(dr.x)); md[1] = fmaxf(md[1], fabsf(dr.y));
            md[2] = fmaxf(md[2], fabsf(dr.z)); md[3] = fmaxf(md[3], fabsf(dr.w));
            mx[0] = fmaxf(mx[0], fabsf(xh.x)); mx[1] = fmaxf(mx[1], fabsf(xh.y));
            mx[2] = fmaxf(mx[2], fabsf(xh.z)); mx[3] = fmaxf(mx[3], fabsf(xh.w));
        }
    }
    __shared__ float red[8][16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        for (int o = 16; o > 0; o >>= 1) {
            s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
            s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
            md[j] = fmaxf(md[j], __shfl_xor_sync(0xffffffffu, md[j], o));
            mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], o));
        }
    }
    if (lane == 0) {
        for (int j = 0; j < 4; j++) { red[warp][j] = s1[j]; red[warp][4 + j] = s2[j]; red[warp][8 + j] = md[j]; red[warp][12 + j] = mx[j]; }
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += (double)red[w][threadIdx.x];
        const int j = threadIdx.x & 3, which = threadIdx.x >> 2;
        atomicAdd(p.sums + (nc + j) * 2 + which, t);
    } else if (threadIdx.x < 16 && p.amax) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t = fmaxf(t, red[w][threadIdx.x]);
        const int j = threadIdx.x & 3, which = (threadIdx.x >> 2) - 2;
        if (!(t < 3.0e38f)) t = 3.0e38f;                     // inf / nan: the scale falls back to 1
        atomicMax(p.amax + (nc + j) * 2 + which, __float_as_uint(t));   // non-negative floats order like their bits
    }
    if (p.slope_sums) {                       // (block-uniform; the shared buffer is free again after this barrier)
        __syncthreads();
        for (int o = 16; o > 0; o >>= 1) s3 += __shfl_xor_sync(0xffffffffu, s3, o);
        if (lane == 0) red[warp][0] = s3;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += (double)red[w][0];
            atomicAdd(p.slope_sums + nc, t);
        }
    }
}

// nn.PReLU: the slope's gradient = the sum of the per-(n, channel quad) partial sums
__global__ void prelu_finalize_kernel(const double* __restrict__ slope_sums, int n, float* __restrict__ dslope)
{
    __shared__ double red[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) t += slope_sums[i];
    red[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) dslope[0] = (float)red[0];
}

__global__ void norm_bwd_finalize_kernel(const double* __restrict__ sums, const double* __restrict__ fwd_stats, int mode,
                                         int G, int N, int C, int Cp, double S, const float* __restrict__ gamma,
                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                         float* __restrict__ m1o, float* __restrict__ m2o, float* __restrict__ dgamma,
                                         float* __restrict__ dbeta, float* __restrict__ dbias,
                                         const unsigned int* __restrict__ amax, float* __restrict__ dy_scale)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * Cp) return;
    const int n = i / Cp, c = i % Cp;
    if (c >= C) { m1o[i] = 0.f; m2o[i] = 0.f; return; }
    double m1 = 0.0, m2 = 0.0;
    if (mode == 1) {
        const int cg = C / G, g = c / cg;
        for (int j = 0; j < cg; j++) {
            const int cc = g * cg + j;
            const double ga = gamma ? (double)gamma[cc] : 1.0;
            m1 += ga * sums[((size_t)n * Cp + cc) * 2];
            m2 += ga * sums[((size_t)n * Cp + cc) * 2 + 1];
        }
        m1 /= S * cg; m2 /= S * cg;
    } else if (mode == 2) {
        const double ga = gamma ? (double)gamma[c] : 1.0;
        for (int j = 0; j < N; j++) { m1 += sums[((size_t)j * Cp + c) * 2]; m2 += sums[((size_t)j * Cp + c) * 2 + 1]; }
        m1 *= ga / (S * N); m2 *= ga / (S * N);
    }
    m1o[i] = (float)m1; m2o[i] = (float)m2;
    if (amax && dy_scale) {
        // |dy| = |rstd (gamma dr - m1 - xhat m2)| <= rstd (|gamma| max|dr| + |m1| + max|xhat| |m2|)
        const float ga = (gamma && mode != 0) ? fabsf(gamma[c]) : 1.f;
        const float r = rstd ? rstd[i] : 1.f;
        const float bound = r * (ga * __uint_as_float(amax[(size_t)i * 2]) + fabsf((float)m1) +
                                 __uint_as_float(amax[(size_t)i * 2 + 1]) * fabsf((float)m2));
        atomicMax(reinterpret_cast<unsigned int*>(dy_scale), __float_as_uint(bound < 3.0e38f ? bound : 3.0e38f));
    }
    if (n != 0) return;
    // per-channel parameter gradients (loop over the batch; for group mode m1/m2 differ per sample)
    double dg = 0.0, db = 0.0, dbi = 0.0;
    for (int j = 0; j < N; j++) {
        const double S1 = sums[((size_t)j * Cp + c) * 2], S2 = sums[((size_t)j * Cp + c) * 2 + 1];
        dg += S2; db += S1;
        if (dbias) {
            const double ga = (gamma && mode != 0) ? (double)gamma[c] : 1.0;
            const double r = (double)rstd[(size_t)j * Cp + c], mu = (double)mean[(size_t)j * Cp + c];
            double mm1 = 0.0, mm2 = 0.0;
            if (mode == 1) {
                const int cg = C / G, g = c / cg;
                for (int q = 0; q < cg; q++) {
                    const int cc = g * cg + q;
                    const double gq = gamma ? (double)gamma[cc] : 1.0;
                    mm1 += gq * sums[((size_t)j * Cp + cc) * 2];
                    mm2 += gq * sums[((size_t)j * Cp + cc) * 2 + 1];
                }
                mm1 /= S * cg; mm2 /= S * cg;
            } else if (mode == 2) { mm1 = m1; mm2 = m2; }
            double sum_xhat = 0.0;
            if (mode == 1 || mode == 2) sum_xhat = r * (fwd_stats[((size_t)j * C + c) * 2] - S * mu);
            dbi += r * (ga * S1 - S * mm1 - mm2 * sum_xhat);
        }
    }
    if (dgamma) dgamma[c] = (float)dg;
    if (dbeta) dbeta[c] = (float)db;
    if (dbias) dbias[c] = (float)dbi;
}

// The same finalisation for problems that fit ONE block (N * pad8(C) <= 1024, i.e. every BASELINE layer): thread
// (n, c) computes the means of its sample and leaves its share of the parameter gradients in shared memory; the
// threads of sample 0 then add the shares in sample order (deterministic).  The general kernel above spends ~8 us
// in the serial N x (C/G) fp64 loops of the sample-0 threads; this one ~3 us.
__global__ void __launch_bounds__(1024) norm_bwd_finalize_block_kernel(
    const double* __restrict__ sums, const double* __restrict__ fwd_stats, int mode, int G, int N, int C, int Cp, double S,
    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ m1o,
    float* __restrict__ m2o, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
    const unsigned int* __restrict__ amax, float* __restrict__ dy_scale)
{
    extern __shared__ double share[];            // [N * Cp][3]: S2 (dgamma), S1 (dbeta), dbias share of sample n
    __shared__ unsigned int bound_bits;
    const int i = threadIdx.x;
    if (i == 0) bound_bits = 0u;
    __syncthreads();
    const bool on = i < N * Cp;
    const int n = on ? i / Cp : 0, c = on ? i % Cp : 0;
    double m1 = 0.0, m2 = 0.0;
    if (on && c < C) {
        if (mode == 1) {
            const int cg = C / G, g = c / cg;
            for (int j = 0; j < cg; j++) {
                const int cc = g * cg + j;
                const double ga = gamma ? (double)gamma[cc] : 1.0;
                m1 += ga * sums[((size_t)n * Cp + cc) * 2];
                m2 += ga * sums[((size_t)n * Cp + cc) * 2 + 1];
            }
            m1 /= S * cg; m2 /= S * cg;
        } else if (mode == 2) {
            const double ga = gamma ? (double)gamma[c] : 1.0;
            for (int j = 0; j < N; j++) { m1 += sums[((size_t)j * Cp + c) * 2]; m2 += sums[((size_t)j * Cp + c) * 2 + 1]; }
            m1 *= ga / (S * N); m2 *= ga / (S * N);
        }
        const double S1 = sums[(size_t)i * 2], S2 = sums[(size_t)i * 2 + 1];
        double dbi = 0.0;
        if (dbias) {
            const double ga = (gamma && mode != 0) ? (double)gamma[c] : 1.0;
            const double r = (double)rstd[i], mu = (double)mean[i];
            double sum_xhat = 0.0;
            if (mode == 1 || mode == 2) sum_xhat = r * (fwd_stats[((size_t)n * C + c) * 2] - S * mu);
            dbi = r * (ga * S1 - S * m1 - m2 * sum_xhat);
        }
        share[i * 3] = S2; share[i * 3 + 1] = S1; share[i * 3 + 2] = dbi;
        if (amax && dy_scale) {
            // |dy| = |rstd (gamma dr - m1 - xhat m2)| <= rstd (|gamma| max|dr| + |m1| + max|xhat| |m2|)
            const float ga = (gamma && mode != 0) ? fabsf(gamma[c]) : 1.f;
            const float r = rstd ? rstd[i] : 1.f;
            const float bound = r * (ga * __uint_as_float(amax[(size_t)i * 2]) + fabsf((float)m1) +
                                     __uint_as_float(amax[(size_t)i * 2 + 1]) * fabsf((float)m2));
            atomicMax(&bound_bits, __float_as_uint(bound < 3.0e38f ? bound : 3.0e38f));
        }
    }
    if (on) { m1o[i] = (float)m1; m2o[i] = (float)m2; }
    __syncthreads();
    if (i == 0 && amax && dy_scale) atomicMax(reinterpret_cast<unsigned int*>(dy_scale), bound_bits);
    if (on && n == 0 && c < C) {
        double dg = 0.0, db = 0.0, dbi = 0.0;
        for (int j = 0; j < N; j++) {
            const double* sh = share + ((size_t)j * Cp + c) * 3;
            dg += sh[0]; db += sh[1]; dbi += sh[2];
        }
        if (dgamma) dgamma[c] = (float)dg;
        if (dbeta) dbeta[c] = (float)db;
        if (dbias) dbias[c] = (float)dbi;
    }
}

// grid: (chunks, Cq, N) over the flat thread grid (Dg,Hg,Wg) = (D,H,W), or the un-cropped fine grid for
// s2d output; kAppVpt voxels per thread, all loads issued before the first use.
static constexpr int kAppVpt = 1;
__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(const NormBwdDev p)
{
    // blocks walk the tensor BACKWARDS: the reduce pass that ran just before read it forwards, its tail is still in L2
    const int cq = gridDim.y - 1 - blockIdx.y, n = gridDim.z - 1 - blockIdx.z;
    const int bx = gridDim.x - 1 - blockIdx.x;
    const size_t nc = ((size_t)n * p.Cq + cq) * 4;
    float4 mu, rs, sc, sh, m1, m2;
    load_nc4(p.mean, nc, mu, 0.f); load_nc4(p.rstd, nc, rs, 1.f);
    load_nc4(p.scale, nc, sc, 1.f); load_nc4(p.shift, nc, sh, 0.f);
    load_nc4(p.m1, nc, m1, 0.f); load_nc4(p.m2, nc, m2, 0.f);
    float4 ga = make_float4(1.f, 1.f, 1.f, 1.f);
    if (p.gamma) {
        const int c = cq * 4;
        ga.x = c < p.C ? p.gamma[c] : 0.f; ga.y = c + 1 < p.C ? p.gamma[c + 1] : 0.f;
        ga.z = c + 2 < p.C ? p.gamma[c + 2] : 0.f; ga.w = c + 3 < p.C ? p.gamma[c + 3] : 0.f;
    }
    const int HWg = p.Hg * p.Wg, Sg = p.Dg * HWg;
    const int v0 = bx * (256 * kAppVpt) + threadIdx.x;
    // fp16 range: dy is stored multiplied by 2^k (undone by the dgrad epilogue through dy_scale[2])
    const float dscale = dy_scale_from_bound(p.dy_scale[0]);
    if (bx == 0 && cq == 0 && n == 0 && threadIdx.x == 0) { p.dy_scale[1] = dscale; p.dy_scale[2] = 1.f / dscale; }
    VoxIn in[kAppVpt];
    int zs[kAppVpt], ys[kAppVpt], xs[kAppVpt];
#pragma unroll
    for (int j = 0; j < kAppVpt; j++) {
        const int v = v0 + j * 256;
        zs[j] = v / HWg;
        const int hw = v - zs[j] * HWg;
        ys[j] = hw / p.Wg; xs[j] = hw - ys[j] * p.Wg;
        if (v < Sg && zs[j] < p.D && ys[j] < p.H && xs[j] < p.W) {
            const size_t ov = ((((size_t)n * p.Cq + cq) * p.D + zs[j]) * p.H + ys[j]) * p.W + xs[j];
            load_vox<true>(p, ov, n, cq, zs[j], ys[j], xs[j], in[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < kAppVpt; j++) {
        const int v = v0 + j * 256;
        if (v >= Sg) continue;
        const int z = zs[j], yy = ys[j], x = xs[j];
        const bool inside = z < p.D && yy < p.H && x < p.W;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inside) {
            float4 dr, xh;
            voxel_grad(p, in[j], mu, rs, sc, sh, dr, xh, p.slope_dev ? __ldg(p.slope_dev) : p.slope);
            o.x = rs.x * (ga.x * dr.x - m1.x - xh.x * m2.x);
            o.y = rs.y * (ga.y * dr.y - m1.y - xh.y * m2.y);
            o.z = rs.z * (ga.z * dr.z - m1.z - xh.z * m2.z);
            o.w = rs.w * (ga.w * dr.w - m1.w - xh.w * m2.w);
            // dy is the MMA operand of dgrad and wgrad: store it rounded to TF32
            o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w);
            if (!p.s2d) {
                const float4 os = make_float4(o.x * dscale, o.y * dscale, o.z * dscale, o.w * dscale);
                store_qh(p.dy, os, n, p.Ch, cq, (size_t)p.D * p.H * p.W, ((size_t)z * p.H + yy) * p.W + x);
            }
        }
        if (p.s2d) {
            // space-to-depth: channel = slot*Cp + c on the coarse grid; fine voxels cropped away by autocrop -> 0
            const int zw = z / p.wd, yw = yy / p.wh, xw = x / p.ww;
            const int slot = ((z - zw * p.wd) * p.wh + (yy - yw * p.wh)) * p.ww + (x - xw * p.ww);
            const size_t wins = (size_t)p.Dw * p.Hw * p.Ww;
            const size_t iw = ((size_t)zw * p.Hw + yw) * p.Ww + xw;
            const float4 os = make_float4(o.x * dscale, o.y * dscale, o.z * dscale, o.w * dscale);
            store_qh(p.dy, os, n, p.Ch, slot * p.Cq + cq, wins, iw);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Vector ("x4") variant for W % 4 == 0: one thread owns 4 consecutive x voxels of one row, so the index arithmetic is
// paid once per 4 voxels (the scalar kernel above is instruction-issue bound, not HBM bound: ~27 issue slots per voxel
// quad are available at full HBM rate).
// ------------------------------------------------------------------------------------------------
E3B_DEVINL float4 apply_formula(const float4& dr, const float4& xh, const float4& rs, const float4& ga, const float4& m1,
                                const float4& m2)
{
    float4 o;
    o.x = tf32_rn(rs.x * (ga.x * dr.x - m1.x - xh.x * m2.x));
    o.y = tf32_rn(rs.y * (ga.y * dr.y - m1.y - xh.y * m2.y));
    o.z = tf32_rn(rs.z * (ga.z * dr.z - m1.z - xh.z * m2.z));
    o.w = tf32_rn(rs.w * (ga.w * dr.w - m1.w - xh.w * m2.w));
    return o;
}

struct BwdConsts { float4 mu, rs, sc, sh, m1, m2, ga; };

E3B_DEVINL void load_bwd_consts(const NormBwdDev& p, int n, int cq, BwdConsts& c, bool with_m)
{
    const size_t nc = ((size_t)n * p.Cq + cq) * 4;
    load_nc4(p.mean, nc, c.mu, 0.f); load_nc4(p.rstd, nc, c.rs, 1.f);
    load_nc4(p.scale, nc, c.sc, 1.f); load_nc4(p.shift, nc, c.sh, 0.f);
    c.m1 = c.m2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (with_m) { load_nc4(p.m1, nc, c.m1, 0.f); load_nc4(p.m2, nc, c.m2, 0.f); }
    c.ga = make_float4(1.f, 1.f, 1.f, 1.f);
    if (p.gamma) {
        const int ch = cq * 4;
        c.ga.x = ch < p.C ? p.gamma[ch] : 0.f; c.ga.y = ch + 1 < p.C ? p.gamma[ch + 1] : 0.f;
        c.ga.z = ch + 2 < p.C ? p.gamma[ch + 2] : 0.f; c.ga.w = ch + 3 < p.C ? p.gamma[ch + 3] : 0.f;
    }
}

// grid: (chunks of 256 x-groups, Cq, N)
__global__ void __launch_bounds__(256) norm_bwd_apply_x4_kernel(const NormBwdDev p)
{
    // blocks walk the tensor BACKWARDS: the reduce pass that ran just before read it forwards, its tail is still in L2
    const int cq = gridDim.y - 1 - blockIdx.y, n = gridDim.z - 1 - blockIdx.z;
    const int bx = gridDim.x - 1 - blockIdx.x;
    BwdConsts c;
    load_bwd_consts(p, n, cq, c, true);
    const int Wg = p.W >> 2;
    const int total = p.D * p.H * Wg;
    const int t = bx * 256 + threadIdx.x;
    const float dscale = dy_scale_from_bound(p.dy_scale[0]);
    if (bx == 0 && cq == 0 && n == 0 && threadIdx.x == 0) { p.dy_scale[1] = dscale; p.dy_scale[2] = 1.f / dscale; }
    if (t >= total) return;
    const int row = t / Wg;
    const int xg = t - row * Wg;
    const int z = row / p.H, yy = row - z * p.H, x0 = xg * 4;
    const size_t ov = (((size_t)n * p.Cq + cq) * p.D * p.H + row) * (size_t)p.W + x0;
    VoxIn in[4];
#pragma unroll
    for (int j = 0; j < 4; j++) load_vox<true>(p, ov + j, n, cq, z, yy, x0 + j, in[j]);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float4 dr, xh;
        voxel_grad(p, in[j], c.mu, c.rs, c.sc, c.sh, dr, xh, p.slope_dev ? __ldg(p.slope_dev) : p.slope);
        const float4 o = apply_formula(dr, xh, c.rs, c.ga, c.m1, c.m2);
        const float4 os = make_float4(o.x * dscale, o.y * dscale, o.z * dscale, o.w * dscale);
        store_qh(p.dy, os, n, p.Ch, cq, (size_t)p.D * p.H * p.W, (size_t)row * p.W + x0 + j);
    }
}

// ------------------------------------------------------------------------------------------------
// Fused norm backward: reduce -> finalize -> apply in ONE persistent kernel, one sample at a time.
//
// The three-kernel path above reads y and the incoming gradient twice from HBM (603 MB per full-resolution 32-channel
// layer of BASELINE cfg 2) with register-staged loads whose memory parallelism is bounded by occupancy.  Here:
//   * group / instance normalisation has per-SAMPLE statistics and one sample's y + gradient (67 MB) fits the 126 MB L2:
//     the kernel makes one ROUND per sample -- phase A reduces the sample, a grid barrier, phase B derives the group means
//     (O(C) work, every CTA redundantly), phase C applies the backward formula to the same items.  y and the gradients
//     cross HBM once; the second read is served by L2 -- or by shared memory, see below;
//   * every CTA owns a fixed contiguous range of work items (an item = 512 voxels x 8 channels of every source tensor)
//     and moves them through a ring of shared-memory stages with 1D bulk copies (cp.async.bulk + mbarrier): the copies
//     of the next stages are in flight while the CTA computes, independent of register count and occupancy;
//   * phase C walks the items BACKWARDS through the same ring: the last S items of phase A are still in shared memory and
//     are not re-read at all (3 of ~7 items per CTA and round for the full-resolution layers of cfg 2), older items come
//     out of L2 newest first; stages freed at the tail of phase C already prefetch the next sample's first items;
//   * dy leaves as full 16-byte units (both 4-channel halves written by one thread).
// Batch normalisation (statistics over the batch) runs the same code as ONE round over all samples.  The gradient's
// power-of-two fp16 scale stays per tensor: it is taken from the first sample's bound; if a later sample needs a smaller
// one, the samples already written are re-scaled in place (exact: a power of two).
// ------------------------------------------------------------------------------------------------
// q = v / d for 0 <= v < 2^31 as one wide multiply and a shift: mul = ceil(2^(31+l) / d), l = ceil(log2 d)
// (error term < 1/d: exact)
struct FastDiv {
    uint32_t mul, shift;
    E3B_DEVINL int div(int v) const { return (int)(((unsigned long long)(unsigned)v * mul) >> shift); }
};
static FastDiv make_fastdiv(int d)
{
    int l = 0;
    while ((1ll << l) < d) l++;
    FastDiv f;
    f.shift = 31 + l;
    f.mul = (uint32_t)(((1ull << f.shift) + (unsigned)d - 1) / (unsigned)d);
    return f;
}

// Optional phase timeline of the fused kernel (E3B_FUSED_PROF, scripts/normbwd_bench.py): globaltimer stamps of thread 0 of
// the first and the last CTA, [cta 2][round 4][stamp 8]: 0 round start, 1 phase A done, 2 grid barrier passed, 3 phase B
// done, 4 phase C done, 5 / 6 time spent waiting for copies in phase A / C.
__device__ unsigned long long g_fused_prof[64];

struct FusedDev {
    int mode, G;
    double S;
    const double* fwd_stats;
    float *dgamma, *dbeta, *dbias;
    unsigned int* counter;               // grid barrier (zeroed by the launch wrapper)
    int per_sample;                      // 1: statistics per sample (group / instance / none), 0: over the batch
    int nset;                            // samples per round: a divisor of N (batch statistics: N)
    int Cp;
    int stages;                          // ring depth
    int g1_bulk;                         // the second gradient is staged by bulk copies (present and not cropped)
    int prof;
    double inv_count;                    // 1 / (elements a mean is taken over): S * C/G (group), S * N (batch)
    FastDiv dHW, dW, dwd, dwh, dww;      // GENERAL variant: voxel index -> (z, y, x) and pooling / s2d window coordinates
    // variant 5: the pooled gradient and the arg-max slots of an item's pooling windows are staged with the item.
    // An item then covers whole windows: co_n coarse voxels starting at coarse index item_chunk_base(chunk) of the slab.
    int co_n;                            // coarse voxels per item (a multiple of 4, <= 128)
    int co_ipp;                          // items per fine plane (co_mode 1), unused otherwise
    int co_mode;                         // 1: an item is 512 / W whole rows of one plane; 2: an item is 512 / (H W) whole planes
};

// first coarse voxel (index inside the (n, quad) slab of the pooled tensor) of the windows item `chunk` covers
E3B_DEVINL int fused_coarse_base(const NormBwdDev& p, const FusedDev& f, int chunk)
{
    if (f.co_mode == 2) return chunk * f.co_n;
    const int z = chunk / f.co_ipp, y0 = (chunk - z * f.co_ipp) * (512 / p.W);
    return ((z / p.wd) * p.Hw + y0 / p.wh) * p.Ww;
}

static constexpr int kFusedMaxCp = 512;
static constexpr int kItemVox = 512;                         // voxels per work item (x 2 channel quads x 16 B = 16 KB per tensor)
static constexpr int kItemTensorBytes = 2 * kItemVox * 16;
static constexpr int kFusedMaxStages = 8;
static constexpr int kCtabSlots = kFusedMaxCp / 4;          // the constants of ALL rounds are staged up front when they fit

// mu holds -mean * rstd: xhat = fma(y, rstd, mu).  slope: what the gradient is multiplied by where fma(y, sc, sh) <= 0
// (0 ReLU, 1 no activation, else the leaky-ReLU slope)
struct QuadConsts { float4 mu, rs, sc, sh; float slope; };

static constexpr int kQuadsPerThread = kItemVox / 128;       // a group of 4 warps covers the item's 512 voxels of one channel quad

// 16-byte shared-memory load from a 32-bit shared address (volatile: stays behind the mbarrier wait that guards the stage)
E3B_DEVINL float4 lds128(uint32_t saddr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}

// Kernel variants: what is present is a compile-time fact wherever the common layer types allow it (no dead code, no
// registers for absent inputs); variant 4 decides everything at run time.
//   0: g0                         (conv -> norm -> relu inside a block)
//   1: g0 + un-cropped g1         (last conv of an encoder block at the bottom level / SAME skip + direct gradient)
//   2: g0, space-to-depth output  (norm0 of an UpConv: dy feeds the transposed conv's GEMMs)
//   3: un-cropped g1 + pooled gp  (last conv of an encoder block: skip gradient + un-pooling)
//   4: anything else (cropped skip gradient of VALID nets, ...)
//   5: as 3, with the pooled gradient and the arg-max slots staged in shared memory with the item (items that cover whole
//      pooling windows: even extents, rows / planes that divide the item) instead of gathered per voxel through L2
template <int VAR>
struct FusedVar {
    static constexpr bool coords = VAR >= 2;       // voxel coordinates are needed at all
    static constexpr bool gp_staged = VAR == 5;
    E3B_DEVINL static bool g0(const NormBwdDev& p) { return VAR == 4 ? p.g0 != nullptr : (VAR != 3 && VAR != 5); }
    E3B_DEVINL static bool g1_bulk(const FusedDev& f) { return VAR == 4 ? f.g1_bulk != 0 : (VAR == 1 || VAR == 3 || VAR == 5); }
    E3B_DEVINL static bool g1_crop(const NormBwdDev& p, const FusedDev& f) { return VAR == 4 && p.g1 != nullptr && !f.g1_bulk; }
    E3B_DEVINL static bool gp(const NormBwdDev& p) { return VAR == 4 ? p.gp != nullptr : (VAR == 3 || VAR == 5); }
    E3B_DEVINL static bool s2d(const NormBwdDev& p) { return VAR == 4 ? p.s2d != 0 : VAR == 2; }
};

// where an item's staged pooling windows sit in shared memory (this thread's channel quad) and which coarse voxel is first
// (variant 5: an item is aligned to the pooling windows, so which coarse voxel and which window slot a thread's k-th voxel
// of an item belongs to is the same for every item: tab[k] = local coarse index << 8 | slot inside the plane pair; zslot adds
// the item's z parity)
struct CoarseStage { uint32_t gp, idx; int base; int zslot; int tab[4]; };

E3B_DEVINL void fused_unpool_staged(const CoarseStage& co, int k, float4& g)
{
    const uint32_t e = (uint32_t)co.tab[k];
    const uint32_t local = e >> 8, slot = (e & 255u) + (uint32_t)co.zslot;
    uint32_t raw;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(raw) : "r"(co.idx + local * 4u));
    const float4 t = lds128(co.gp + local * 16u);
    if ((raw & 255u) == slot) g.x += t.x;
    if (((raw >> 8) & 255u) == slot) g.y += t.y;
    if (((raw >> 16) & 255u) == slot) g.z += t.z;
    if ((raw >> 24) == slot) g.w += t.w;
}

// y and the staged gradients (g0 [+ g1]) of this thread's voxel quads of one item: all loads are issued before the first use
static constexpr int kQuadBatch = 2;                          // quads whose loads are in flight together (register budget: 96)

template <int VAR, bool FULL>
E3B_DEVINL void fused_load_item(const NormBwdDev& p, const FusedDev& f, uint32_t sbase, uint32_t off_g0, uint32_t off_g1, int t128,
                                int nv, int k0, float4* Y, float4* G)
{
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kk = 0; kk < kQuadBatch; kk++) {
        const int k = k0 + kk;
        Y[kk] = zero; G[kk] = zero;
        if (!FULL && k * 128 + t128 >= nv) continue;
        Y[kk] = lds128(sbase + k * 2048);
        if (FusedVar<VAR>::g0(p)) G[kk] = lds128(sbase + off_g0 + k * 2048);
    }
    if (FusedVar<VAR>::g1_bulk(f)) {
#pragma unroll
        for (int kk = 0; kk < kQuadBatch; kk++) {
            const int k = k0 + kk;
            if (!FULL && k * 128 + t128 >= nv) continue;
            const float4 t = lds128(sbase + off_g1 + k * 2048);
            G[kk].x += t.x; G[kk].y += t.y; G[kk].z += t.z; G[kk].w += t.w;
        }
    }
}

// voxel coordinates, and the gradients that are gathered from global memory (cropped skip gradient, un-pooling of the
// pooled gradient; L1 / L2 hits: 8 fine voxels share one coarse voxel)
template <int VAR>
E3B_DEVINL void fused_gather(const NormBwdDev& p, const FusedDev& f, int n, int cq, int v, float4& g, int& z, int& yy, int& x,
                             const CoarseStage& co)
{
    z = f.dHW.div(v);
    const int r = v - z * (p.H * p.W);
    yy = f.dW.div(r); x = r - yy * p.W;
    if (FusedVar<VAR>::g1_crop(p, f)) {
        // backward of autocrop's slice of the skip tensor: zero outside the cropped box
        const int zc = z - p.g1_od, yc = yy - p.g1_oh, xc = x - p.g1_ow;
        if (zc >= 0 && zc < p.g1_D && yc >= 0 && yc < p.g1_H && xc >= 0 && xc < p.g1_W) {
            const float4 t = __ldg(p.g1 + ((((size_t)n * p.Cq + cq) * p.g1_D + zc) * p.g1_H + yc) * p.g1_W + xc);
            g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
    }
    if (FusedVar<VAR>::gp(p)) {
        const int zw = f.dwd.div(z), yw = f.dwh.div(yy), xw = f.dww.div(x);
        const unsigned char slot = (unsigned char)(((z - zw * p.wd) * p.wh + (yy - yw * p.wh)) * p.ww + (x - xw * p.ww));
        uchar4 idxs;
        float4 t;
        if (FusedVar<VAR>::gp_staged) {
            const uint32_t local = (uint32_t)(((zw * p.Hw + yw) * p.Ww + xw) - co.base);
            uint32_t raw;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(raw) : "r"(co.idx + local * 4u));
            idxs = make_uchar4((unsigned char)(raw & 255u), (unsigned char)((raw >> 8) & 255u), (unsigned char)((raw >> 16) & 255u),
                               (unsigned char)(raw >> 24));
            t = lds128(co.gp + local * 16u);
        } else {
            const size_t ow = ((((size_t)n * p.Cq + cq) * p.Dw + zw) * p.Hw + yw) * p.Ww + xw;
            idxs = p.pool_idx[ow];
            t = p.gp[ow];
        }
        if (idxs.x == slot) g.x += t.x;
        if (idxs.y == slot) g.y += t.y;
        if (idxs.z == slot) g.z += t.z;
        if (idxs.w == slot) g.w += t.w;
    }
}

// xhat, and the (leaky-)ReLU derivative applied to the summed upstream gradient: 1 where y * scale + shift > 0 (scale = 1,
// shift = 0 without a norm), the negative slope elsewhere.  The forward's TF32 rounding of the activation cannot turn a
// positive value into zero, so the mask does not need it.
template <bool SILU>
E3B_DEVINL void fused_mask_xhat(const QuadConsts& c, const float4& yv, float4& g, float4& xh)
{
    xh = make_float4(fmaf(yv.x, c.rs.x, c.mu.x), fmaf(yv.y, c.rs.y, c.mu.y), fmaf(yv.z, c.rs.z, c.mu.z), fmaf(yv.w, c.rs.w, c.mu.w));
    if (SILU) {
        // nn.SiLU (a compile-time variant of the kernel: the ReLU family pays nothing for it)
        g.x = act_bwd(fmaf(yv.x, c.sc.x, c.sh.x), g.x, 2, 0.f); g.y = act_bwd(fmaf(yv.y, c.sc.y, c.sh.y), g.y, 2, 0.f);
        g.z = act_bwd(fmaf(yv.z, c.sc.z, c.sh.z), g.z, 2, 0.f); g.w = act_bwd(fmaf(yv.w, c.sc.w, c.sh.w), g.w, 2, 0.f);
        return;
    }
    if (!(fmaf(yv.x, c.sc.x, c.sh.x) > 0.f)) g.x *= c.slope;
    if (!(fmaf(yv.y, c.sc.y, c.sh.y) > 0.f)) g.y *= c.slope;
    if (!(fmaf(yv.z, c.sc.z, c.sh.z) > 0.f)) g.z *= c.slope;
    if (!(fmaf(yv.w, c.sc.w, c.sh.w) > 0.f)) g.w *= c.slope;
}

// phase A over one staged item (this thread: channel quad `cq`, voxels t128 + 128 k).
// FULL = all 512 voxels present: no bounds checks.
template <int VAR, bool FULL, bool SILU>
E3B_DEVINL void fused_item_reduce(const NormBwdDev& p, const FusedDev& f, uint32_t sbase, uint32_t off_g0, uint32_t off_g1,
                                  const QuadConsts& c, int t128, int n, int cq, int v0, int nv, float* s1, float* s2, float* md,
                                  float* mx, const CoarseStage& co)
{
#pragma unroll
    for (int k0 = 0; k0 < kQuadsPerThread; k0 += kQuadBatch) {
    float4 Y[kQuadBatch], G[kQuadBatch];
    fused_load_item<VAR, FULL>(p, f, sbase, off_g0, off_g1, t128, nv, k0, Y, G);
#pragma unroll
    for (int kk = 0; kk < kQuadBatch; kk++) {
        const int vl = (k0 + kk) * 128 + t128;
        if (!FULL && vl >= nv) continue;
        float4 dr = G[kk], xh;
        if (FusedVar<VAR>::gp_staged) fused_unpool_staged(co, k0 + kk, dr);
        else if (FusedVar<VAR>::coords) { int z, yy, x; fused_gather<VAR>(p, f, n, cq, v0 + vl, dr, z, yy, x, co); }
        fused_mask_xhat<SILU>(c, Y[kk], dr, xh);
        s1[0] += dr.x; s1[1] += dr.y; s1[2] += dr.z; s1[3] += dr.w;
        s2[0] = fmaf(dr.x, xh.x, s2[0]); s2[1] = fmaf(dr.y, xh.y, s2[1]); s2[2] = fmaf(dr.z, xh.z, s2[2]); s2[3] = fmaf(dr.w, xh.w, s2[3]);
        md[0] = fmaxf(md[0], fabsf(dr.x)); md[1] = fmaxf(md[1], fabsf(dr.y)); md[2] = fmaxf(md[2], fabsf(dr.z)); md[3] = fmaxf(md[3], fabsf(dr.w));
        mx[0] = fmaxf(mx[0], fabsf(xh.x)); mx[1] = fmaxf(mx[1], fabsf(xh.y)); mx[2] = fmaxf(mx[2], fabsf(xh.z)); mx[3] = fmaxf(mx[3], fabsf(xh.w));
    }
    }
}

// phase C over one staged item: dy = 2^k * rstd * (gamma * dr - m1 - xhat * m2); this thread writes the 8-byte half `hsel`
// of its voxels' 16-byte units (the other half comes from the other warp group; L2 merges the sectors).
// (fp16 has TF32's 10 mantissa bits: the fp16 rounding of the store is the operand rounding.)
template <int VAR, bool FULL, bool SILU>
E3B_DEVINL void fused_item_apply(const NormBwdDev& p, const FusedDev& f, uint32_t sbase, uint32_t off_g0, uint32_t off_g1,
                                 const QuadConsts& c, const float4& ga, const float4& m1, const float4& m2, const float4& rk, int hsel,
                                 int t128, int n, int cqp, int Cqp, int total, int v0, int nv, uint2* dy, const CoarseStage& co)
{
    const int cq = 2 * cqp + hsel;
    // this thread's 8-byte half of its first voxel's unit; the other quads follow at 128 voxels = 2 KB
    uint2* const dyp = dy + ((((size_t)n * p.Ch + cqp) * (size_t)total + (size_t)(v0 + t128)) * 2 + hsel);
#pragma unroll
    for (int k0 = 0; k0 < kQuadsPerThread; k0 += kQuadBatch) {
    float4 Y[kQuadBatch], G[kQuadBatch];
    fused_load_item<VAR, FULL>(p, f, sbase, off_g0, off_g1, t128, nv, k0, Y, G);
#pragma unroll
    for (int kk = 0; kk < kQuadBatch; kk++) {
        const int vl = (k0 + kk) * 128 + t128;
        if (!FULL && vl >= nv) continue;
        float4 dr = G[kk], xh;
        int z = 0, yy = 0, x = 0;
        if (FusedVar<VAR>::gp_staged) fused_unpool_staged(co, k0 + kk, dr);
        else if (FusedVar<VAR>::coords) fused_gather<VAR>(p, f, n, cq, v0 + vl, dr, z, yy, x, co);
        fused_mask_xhat<SILU>(c, Y[kk], dr, xh);
        // rstd * 2^k is folded into rk: o = rk * (ga * dr - m1 - xh * m2)
        const uint2 o = pack_half4(rk.x * fmaf(-xh.x, m2.x, fmaf(ga.x, dr.x, -m1.x)), rk.y * fmaf(-xh.y, m2.y, fmaf(ga.y, dr.y, -m1.y)),
                                   rk.z * fmaf(-xh.z, m2.z, fmaf(ga.z, dr.z, -m1.z)), rk.w * fmaf(-xh.w, m2.w, fmaf(ga.w, dr.w, -m1.w)));
        if (!FusedVar<VAR>::s2d(p)) {
            dyp[(k0 + kk) * 256] = o;
        } else {
            // space-to-depth: channel = slot * Cp + c on the coarse grid
            const int zw = f.dwd.div(z), yw = f.dwh.div(yy), xw = f.dww.div(x);
            const int slot = ((z - zw * p.wd) * p.wh + (yy - yw * p.wh)) * p.ww + (x - xw * p.ww);
            const size_t unit = ((size_t)n * p.Ch + slot * Cqp + cqp) * ((size_t)p.Dw * p.Hw * p.Ww) + ((size_t)zw * p.Hw + yw) * p.Ww + xw;
            dy[unit * 2 + hsel] = o;
        }
    }
    }
}

static constexpr int kFusedThreads = 288;            // 8 consumer warps + 1 producer warp (one lane issues the bulk copies)
static constexpr int kFusedBar = 1;                  // named barrier of the 256 consumer threads

template <int VAR, bool SILU>
__global__ void __launch_bounds__(kFusedThreads, 2) norm_bwd_fused_kernel(const NormBwdDev p, const FusedDev f)
{
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ uint64_t full[kFusedMaxStages], empty[kFusedMaxStages];
    __shared__ float4 ctab[kCtabSlots][5];           // per (sample, 4-channel quad): -mean * rstd, rstd, scale, shift, gamma
    __shared__ float2 gm[kFusedMaxCp];               // (m1, m2) per (sample, group) -- batch statistics: per channel
    __shared__ float red[8][16];
    __shared__ float bound_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = p.D * p.H * p.W;               // voxels per (n, 4-channel plane) slab
    const int nchunks = (total + kItemVox - 1) / kItemVox;
    const int Cqp = p.Cq >> 1;                       // pairs of channel quads (Cq is even: channels are padded to 8)
    const int nset = f.nset, nrounds = p.N / nset;       // samples per round (a divisor of N); batch statistics: one round
    const int Cp = f.Cp, S = f.stages;
    const bool has_g0 = FusedVar<VAR>::g0(p), g1_bulk = FusedVar<VAR>::g1_bulk(f);
    const int nb = 1 + (has_g0 ? 1 : 0) + (g1_bulk ? 1 : 0);
    // variant 5: behind the fine tensors, per channel quad the item's pooled gradient (16 B per coarse voxel), then its slots (4 B)
    const uint32_t co_gp_bytes = FusedVar<VAR>::gp_staged ? (uint32_t)f.co_n * 16u : 0u, co_idx_bytes = FusedVar<VAR>::gp_staged ? (uint32_t)f.co_n * 4u : 0u;
    const uint32_t off_cgp = nb * kItemTensorBytes, off_cidx = off_cgp + 2 * co_gp_bytes;
    const uint32_t stage_bytes = nb * kItemTensorBytes + 2 * (co_gp_bytes + co_idx_bytes);
    const uint32_t off_g0 = kItemTensorBytes, off_g1 = (has_g0 ? 2 : 1) * kItemTensorBytes;
    const long long items = (long long)nset * Cqp * nchunks;                   // < 2^31 (checked by the launch wrapper)
    const int i0 = (int)(items * blockIdx.x / gridDim.x), i1 = (int)(items * (blockIdx.x + 1) / gridDim.x);
    const int cnt = i1 - i0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        fence_barrier_init();
    }
    __syncthreads();
    if (warp == 8) {
        // ---------------- producer: the whole kernel's copy schedule, throttled by the consumers' releases of the stages.
        // Phase A walks the CTA's items forwards through the ring, phase C backwards through the SAME ring: the last S items
        // of phase A are still staged when phase C starts and are not copied again.
        if (lane != 0) return;
        // start the bulk copies of work item `it` of the round whose first sample is n0 into stage s
        auto issue = [&](int n0, int it, int s) {
            const int sp = it / nchunks, chunk = it - sp * nchunks;
            const int n = n0 + sp / Cqp, cqp = sp - (sp / Cqp) * Cqp;
            const int v0 = chunk * kItemVox;
            const uint32_t bytes = (uint32_t)min(kItemVox, total - v0) * 16u;
            unsigned char* dst = ring + (size_t)s * stage_bytes;
            mbar_arrive_expect_tx(&full[s], 2 * nb * bytes + 2 * (co_gp_bytes + co_idx_bytes));
            const size_t cbase = FusedVar<VAR>::gp_staged ? (size_t)fused_coarse_base(p, f, chunk) : 0;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const size_t off = ((size_t)n * p.Cq + 2 * cqp + h) * (size_t)total + v0;
                bulk_load_1d(dst + h * (kItemVox * 16), p.y + off, bytes, &full[s]);
                if (has_g0) bulk_load_1d(dst + off_g0 + h * (kItemVox * 16), p.g0 + off, bytes, &full[s]);
                if (g1_bulk) bulk_load_1d(dst + off_g1 + h * (kItemVox * 16), p.g1 + off, bytes, &full[s]);
                if (FusedVar<VAR>::gp_staged) {
                    const size_t coff = ((size_t)n * p.Cq + 2 * cqp + h) * ((size_t)p.Dw * p.Hw * p.Ww) + cbase;
                    bulk_load_1d(dst + off_cgp + h * co_gp_bytes, p.gp + coff, co_gp_bytes, &full[s]);
                    bulk_load_1d(dst + off_cidx + h * co_idx_bytes, p.pool_idx + coff, co_idx_bytes, &full[s]);
                }
            }
        };
        uint32_t eparity = 0;                         // bit s: the phase of empty[s] the next refill of stage s waits for
        auto refill = [&](int n0, int it, int s) {
            mbar_wait(&empty[s], (eparity >> s) & 1u);
            eparity ^= 1u << s;
            issue(n0, it, s);
        };
        for (int j = 0; j < S && j < cnt; j++) issue(0, i0 + j, j);
        for (int round = 0; round < nrounds; round++) {
            const int n0 = round * nset;
            for (int j = 0; j + S < cnt; j++) refill(n0, i0 + j + S, j % S);                       // phase A
            for (int j = cnt - 1; j >= 0; j--) {                                                   // phase C
                if (j - S >= 0) refill(n0, i0 + j - S, j % S);                                     // an older item of this round
                else if (round + 1 < nrounds) refill(n0 + nset, i0 + j, j % S);                    // the next round's first items
            }
        }
        return;
    }
    // ---------------- consumers (256 threads; they synchronise among themselves on a named barrier)
    const int hsel = warp >> 2, t128 = threadIdx.x & 127;
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t ring_thread = ring_u32 + (uint32_t)(hsel * kItemVox + t128) * 16u;     // this thread's first voxel quad of stage 0
    const float act_slope = p.relu ? p.slope : 1.f;
    int co_tab[4] = {0, 0, 0, 0};
    if (FusedVar<VAR>::gp_staged) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int vl = k * 128 + t128;
            int pz = 0, r = vl;
            if (f.co_mode == 2) { pz = vl / (p.H * p.W); r = vl - pz * (p.H * p.W); }
            const int row = r / p.W, x = r - row * p.W;
            const int local = ((pz / p.wd) * p.Hw + row / p.wh) * p.Ww + x / p.ww;
            const int slot = ((f.co_mode == 2 ? pz % p.wd : 0) * p.wh + row % p.wh) * p.ww + x % p.ww;
            co_tab[k] = (local << 8) | slot;
        }
    }
    uint32_t parity = 0;                              // bit s: the phase of full[s] the next wait on stage s is for
    float dscale = 0.f;                               // the tensor's fp16 scale so far (0: none yet); identical in every CTA
    float bound_max = 0.f;
    unsigned int bar_target = 0;
    // Per-(sample, channel quad) constants.  Batch statistics are the same for every sample: one row per quad (cs = 0).
    // ctab rows: (n - ctab_n0) * Cq + cq; all samples at once when they fit (the launch wrapper guarantees nset * Cq fits).
    const int cs = f.per_sample ? 1 : 0;
    const bool all_rounds = cs * p.N * p.Cq <= kCtabSlots;
    // group means: indexed by the absolute sample when all samples fit (CTA 0 re-uses them for the parameter gradients
    // after the last round), else by the sample's position inside the round
    const bool gm_keep = f.mode != 1 || p.N * f.G <= kFusedMaxCp;
    auto fill_ctab = [&](int first, int count) {
        for (int t = threadIdx.x; t < count * p.Cq; t += 256) {
            const int j = t / p.Cq, cq = t - j * p.Cq;
            const size_t nc = ((size_t)(first + j) * p.Cq + cq) * 4;
            float4* row = ctab[t];
            float4 mu, rs;
            load_nc4(p.mean, nc, mu, 0.f); load_nc4(p.rstd, nc, rs, 1.f);
            row[0] = make_float4(-mu.x * rs.x, -mu.y * rs.y, -mu.z * rs.z, -mu.w * rs.w); row[1] = rs;
            load_nc4(p.scale, nc, row[2], 1.f); load_nc4(p.shift, nc, row[3], 0.f);
            float4 ga = make_float4(1.f, 1.f, 1.f, 1.f);
            if (p.gamma) {
                const int ch = cq * 4;
                ga.x = ch < p.C ? p.gamma[ch] : 0.f; ga.y = ch + 1 < p.C ? p.gamma[ch + 1] : 0.f;
                ga.z = ch + 2 < p.C ? p.gamma[ch + 2] : 0.f; ga.w = ch + 3 < p.C ? p.gamma[ch + 3] : 0.f;
            }
            row[4] = ga;
        }
    };
    if (all_rounds) {
        fill_ctab(0, cs ? p.N : 1);                   // while the first bulk copies are in flight
        named_bar_sync(kFusedBar, 256);
    }
    for (int round = 0; round < nrounds; round++) {
        const int n0 = round * nset;
        const bool prof_on = f.prof && threadIdx.x == 0 && round < 4 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
        unsigned long long* prof = g_fused_prof + ((blockIdx.x == 0 ? 0 : 1) * 4 + (round & 3)) * 8;
        unsigned long long wait_a = 0, wait_c = 0;
        if (prof_on) prof[0] = globaltimer_ns();
        if (!all_rounds) {
            fill_ctab(n0, nset);                      // too many (sample, channel) pairs for the table: one round at a time
            named_bar_sync(kFusedBar, 256);
        }
        const int ctab_n0 = all_rounds ? 0 : n0;
        // ---------------- phase A: sums of dr and dr * xhat, max |dr| and |xhat| per (n, c).
        // Warps 0-3 work on the first channel quad of the item's pair, warps 4-7 on the second: the per-channel constants
        // stay in registers while the slab pair does not change.
        {
            float s1[4], s2[4], md[4], mx[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { s1[j] = 0.f; s2[j] = 0.f; md[j] = 0.f; mx[j] = 0.f; }
            int cur_sp = -1, n = 0, cqp = 0;
            int sp = i0 / nchunks, chunk = i0 - sp * nchunks;                   // item i0 + j = (slab pair sp, chunk)
            QuadConsts c;
            c.slope = act_slope;
            for (int j = 0; j <= cnt; j++) {
                if (j == cnt || sp != cur_sp) {
                    if (cur_sp >= 0) {
                        // this slab pair's partial sums: warp shuffles -> shared memory -> 16 fp64 + 16 max atomics per CTA
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            for (int o = 16; o > 0; o >>= 1) {
                                s1[q] += __shfl_xor_sync(0xffffffffu, s1[q], o);
                                s2[q] += __shfl_xor_sync(0xffffffffu, s2[q], o);
                                md[q] = fmaxf(md[q], __shfl_xor_sync(0xffffffffu, md[q], o));
                                mx[q] = fmaxf(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], o));
                            }
                        }
                        if (lane == 0) {
#pragma unroll
                            for (int q = 0; q < 4; q++) { red[warp][q] = s1[q]; red[warp][4 + q] = s2[q]; red[warp][8 + q] = md[q]; red[warp][12 + q] = mx[q]; }
                        }
                        named_bar_sync(kFusedBar, 256);
                        if (threadIdx.x < 32) {
                            // thread (h, q): entry q of channel quad h, over that quad's four warps
                            const int h = threadIdx.x >> 4, q = threadIdx.x & 15;
                            const size_t ch = ((size_t)n * p.Cq + 2 * cqp + h) * 4 + (q & 3);
                            if (q < 8) {
                                double t = 0.0;
                                for (int w = 0; w < 4; w++) t += (double)red[h * 4 + w][q];
                                atomicAdd(p.sums + ch * 2 + (q >> 2), t);
                            } else {
                                float t = 0.f;
                                for (int w = 0; w < 4; w++) t = fmaxf(t, red[h * 4 + w][q]);
                                if (!(t < 3.0e38f)) t = 3.0e38f;               // inf / nan: the scale falls back to 1
                                atomicMax(p.amax + ch * 2 + ((q >> 2) - 2), __float_as_uint(t));
                            }
                        }
                        named_bar_sync(kFusedBar, 256);
#pragma unroll
                        for (int q = 0; q < 4; q++) { s1[q] = 0.f; s2[q] = 0.f; md[q] = 0.f; mx[q] = 0.f; }
                    }
                    if (j == cnt) break;
                    cur_sp = sp;
                    n = n0 + sp / Cqp; cqp = sp - (sp / Cqp) * Cqp;
                    const float4* row = ctab[cs * (n - ctab_n0) * p.Cq + 2 * cqp + hsel];
                    c.mu = row[0]; c.rs = row[1]; c.sc = row[2]; c.sh = row[3];
                }
                const int v0 = chunk * kItemVox, nv = min(kItemVox, total - v0);
                const int s = j % S;
                const unsigned long long tw = prof_on ? globaltimer_ns() : 0;
                mbar_wait(&full[s], (parity >> s) & 1u);
                if (prof_on) wait_a += globaltimer_ns() - tw;
                parity ^= 1u << s;
                const uint32_t sbase = ring_thread + (uint32_t)s * stage_bytes;
                CoarseStage co;
                co.gp = ring_u32 + (uint32_t)s * stage_bytes + off_cgp + (uint32_t)hsel * co_gp_bytes;
                co.idx = ring_u32 + (uint32_t)s * stage_bytes + off_cidx + (uint32_t)hsel * co_idx_bytes;
                co.base = 0;
                co.zslot = (FusedVar<VAR>::gp_staged && f.co_mode == 1) ? ((chunk / f.co_ipp) % p.wd) * (p.wh * p.ww) : 0;
#pragma unroll
                for (int k = 0; k < 4; k++) co.tab[k] = co_tab[k];
                if (nv == kItemVox) fused_item_reduce<VAR, true, SILU>(p, f, sbase, off_g0, off_g1, c, t128, n, 2 * cqp + hsel, v0, nv, s1, s2, md, mx, co);
                else fused_item_reduce<VAR, false, SILU>(p, f, sbase, off_g0, off_g1, c, t128, n, 2 * cqp + hsel, v0, nv, s1, s2, md, mx, co);
                if (j + S < cnt) {                    // this warp is done with stage s: the producer may refill it
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                }
                if (++chunk == nchunks) { chunk = 0; sp++; }
            }
        }
        // ---------------- grid barrier: every CTA's sums of this round are in
        if (prof_on) { prof[1] = globaltimer_ns(); prof[5] = wait_a; }
        bar_target += gridDim.x;
        named_bar_sync(kFusedBar, 256);
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(f.counter, 1u);
            const uint64_t t0 = globaltimer_ns();
            uint32_t spins = 0;
            while (*reinterpret_cast<volatile unsigned int*>(f.counter) < bar_target) {
                if ((++spins & 0x3FF) == 0 && globaltimer_ns() - t0 > E3B_WAIT_TIMEOUT_NS) __trap();     // never hang the box
            }
            __threadfence();
            bound_s = 0.f;
            if (prof_on) prof[2] = globaltimer_ns();
        }
        named_bar_sync(kFusedBar, 256);
        // ---------------- phase B (every CTA, redundantly): the means m1, m2 per (sample of the round, group) -- batch
        // statistics: per channel -- then the bound on |dy|.  sums / amax were written by other CTAs' atomics: read them
        // from L2 (ld.cg); gamma and rstd come from the shared-memory table.  Two L2 round trips in all: every thread issues
        // its loads before it uses any of them.
        {
            const int nent_c = (cs ? nset : 1) * Cp;                            // (sample, channel) entries of the round
            // this thread's first (sample, channel) entry: its amax loads go out together with the loads of step 1
            const int t0 = threadIdx.x, j0 = t0 / Cp, c0 = t0 - j0 * Cp;
            unsigned int ad0 = 0, ax0 = 0;
            if (t0 < nent_c && c0 < p.C && cs) {
                ad0 = __ldcg(p.amax + ((size_t)(n0 + j0) * Cp + c0) * 2); ax0 = __ldcg(p.amax + ((size_t)(n0 + j0) * Cp + c0) * 2 + 1);
            }
            // step 1: the means.  Group mode with a power-of-two group size: one thread per (sample, channel) loads its own
            // two sums (one L2 round trip for the whole table) and the group's lanes add up by shuffles; otherwise one
            // thread per (sample, group) [group mode] or per channel [batch mode] walks its entries.
            const int cg = f.mode == 1 ? p.C / f.G : 1;
            if (f.mode == 1 && cg <= 32 && (cg & (cg - 1)) == 0 && Cp % cg == 0) {
                for (int t = threadIdx.x; t - lane < nent_c; t += 256) {
                    const int j = t / Cp, c = t - j * Cp, n = n0 + j;
                    const bool valid = t < nent_c && c < p.C;
                    double v1 = 0.0, v2 = 0.0;
                    if (valid) {
                        const double ga = (double)reinterpret_cast<const float*>(&ctab[cs * (n - ctab_n0) * p.Cq + (c >> 2)][4])[c & 3];
                        v1 = ga * __ldcg(p.sums + ((size_t)n * Cp + c) * 2);
                        v2 = ga * __ldcg(p.sums + ((size_t)n * Cp + c) * 2 + 1);
                    }
                    for (int o = cg >> 1; o > 0; o >>= 1) {
                        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
                    }
                    if (valid && (c & (cg - 1)) == 0) gm[(gm_keep ? n : j) * f.G + c / cg] = make_float2((float)(v1 * f.inv_count), (float)(v2 * f.inv_count));
                }
            } else {
                const int ngrp = f.mode == 1 ? f.G : (f.mode == 2 ? p.C : 0);
                for (int t = threadIdx.x; t < (cs ? nset : 1) * ngrp; t += 256) {
                    const int j = t / ngrp, g = t - j * ngrp;
                    double m1 = 0.0, m2 = 0.0;
                    if (f.mode == 1) {
                        const int n = n0 + j;
                        const float4* grow = ctab[cs * (n - ctab_n0) * p.Cq];
#pragma unroll 4
                        for (int q = 0; q < cg; q++) {
                            const int cc = g * cg + q;
                            const double ga = (double)reinterpret_cast<const float*>(&grow[(cc >> 2) * 5 + 4])[cc & 3];
                            m1 += ga * __ldcg(p.sums + ((size_t)n * Cp + cc) * 2);
                            m2 += ga * __ldcg(p.sums + ((size_t)n * Cp + cc) * 2 + 1);
                        }
                    } else {
                        const double ga = (double)reinterpret_cast<const float*>(&ctab[g >> 2][4])[g & 3];
#pragma unroll 4
                        for (int q = 0; q < p.N; q++) {
                            m1 += __ldcg(p.sums + ((size_t)q * Cp + g) * 2);
                            m2 += __ldcg(p.sums + ((size_t)q * Cp + g) * 2 + 1);
                        }
                        m1 *= ga; m2 *= ga;
                    }
                    gm[(f.mode == 1 && gm_keep ? n0 * ngrp : 0) + t] = make_float2((float)(m1 * f.inv_count), (float)(m2 * f.inv_count));
                }
            }
            named_bar_sync(kFusedBar, 256);
            // step 2: |dy| = |rstd (gamma dr - m1 - xhat m2)| <= rstd (|gamma| max|dr| + |m1| + max|xhat| |m2|)
            float bmax = 0.f;
            for (int t = threadIdx.x; t < nent_c; t += 256) {
                const int j = t / Cp, c = t - j * Cp;
                if (c >= p.C) continue;
                const int n = n0 + j;
                const float4* row = ctab[cs * (n - ctab_n0) * p.Cq + (c >> 2)];
                const float ga_abs = f.mode != 0 ? fabsf(reinterpret_cast<const float*>(&row[4])[c & 3]) : 1.f;
                const float rs_c = reinterpret_cast<const float*>(&row[1])[c & 3];
                float2 m = make_float2(0.f, 0.f);
                if (f.mode == 1) m = gm[(gm_keep ? n : j) * f.G + c / (p.C / f.G)];
                else if (f.mode == 2) m = gm[c];
                float b = 0.f;
                for (int q = 0; q < (cs ? 1 : p.N); q++) {
                    const size_t i = (size_t)(n + q) * Cp + c;
                    const unsigned int ad = (cs && t == t0) ? ad0 : __ldcg(p.amax + i * 2), ax = (cs && t == t0) ? ax0 : __ldcg(p.amax + i * 2 + 1);
                    const float bq = rs_c * (ga_abs * __uint_as_float(ad) + fabsf(m.x) + __uint_as_float(ax) * fabsf(m.y));
                    b = fmaxf(b, bq < 3.0e38f ? bq : 3.0e38f);
                }
                bmax = fmaxf(bmax, b);
            }
            // one shared-memory atomic per warp (non-negative floats order like their bits)
            const unsigned int wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(bmax));
            if (lane == 0 && wmax) atomicMax(reinterpret_cast<unsigned int*>(&bound_s), wmax);
        }
        named_bar_sync(kFusedBar, 256);
        const float bound = bound_s;
        if (prof_on) prof[3] = globaltimer_ns();
        bound_max = fmaxf(bound_max, bound);
        const float want = dy_scale_from_bound(bound);
        if (dscale == 0.f) dscale = want;
        else if (want < dscale) {
            // a later round is larger than the scale chosen so far allows: bring the samples already written down to the
            // new scale (a power of two: exact).  Rare -- the samples of one batch have similar gradient magnitudes.
            const float factor = want / dscale;
            const size_t units = (size_t)n0 * p.Ch * (FusedVar<VAR>::s2d(p) ? (size_t)p.Dw * p.Hw * p.Ww : (size_t)total);      // 16-byte units
            uint4* q = reinterpret_cast<uint4*>(p.dy);
            const __half2 f2 = __float2half2_rn(factor);
            for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < units; i += (size_t)gridDim.x * 256) {
                uint4 v = __ldcg(q + i);
                __half2* h = reinterpret_cast<__half2*>(&v);
                h[0] = __hmul2(h[0], f2); h[1] = __hmul2(h[1], f2); h[2] = __hmul2(h[2], f2); h[3] = __hmul2(h[3], f2);
                q[i] = v;
            }
            dscale = want;
        }
        // ---------------- phase C: the backward formula, newest item first
        {
            uint2* dy = reinterpret_cast<uint2*>(p.dy);
            int cur_sp = -1, n = 0, cqp = 0;
            int sp = (i1 - 1) / nchunks, chunk = (i1 - 1) - sp * nchunks;
            QuadConsts c;
            c.slope = act_slope;
            float4 ga, m1, m2, rk;
            for (int j = cnt - 1; j >= 0; j--) {
                if (sp != cur_sp) {
                    cur_sp = sp;
                    n = n0 + sp / Cqp; cqp = sp - (sp / Cqp) * Cqp;
                    const int cq = 2 * cqp + hsel;
                    const float4* row = ctab[cs * (n - ctab_n0) * p.Cq + cq];
                    c.mu = row[0]; c.rs = row[1]; c.sc = row[2]; c.sh = row[3];
                    ga = row[4];
                    {
                        // the means of the four channels' groups
                        float mm[4][2];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int ch = min(4 * cq + k, p.C - 1);
                            float2 m = make_float2(0.f, 0.f);
                            if (f.mode == 1) m = gm[(gm_keep ? n : n - n0) * f.G + ch / (p.C / f.G)];
                            else if (f.mode == 2) m = gm[ch];
                            mm[k][0] = 4 * cq + k < p.C ? m.x : 0.f; mm[k][1] = 4 * cq + k < p.C ? m.y : 0.f;
                        }
                        m1 = make_float4(mm[0][0], mm[1][0], mm[2][0], mm[3][0]);
                        m2 = make_float4(mm[0][1], mm[1][1], mm[2][1], mm[3][1]);
                    }
                    rk = make_float4(c.rs.x * dscale, c.rs.y * dscale, c.rs.z * dscale, c.rs.w * dscale);
                }
                const int v0 = chunk * kItemVox, nv = min(kItemVox, total - v0);
                const int s = j % S;
                if (j < cnt - S) {                    // not one of the items phase A left in the ring
                    const unsigned long long tw = prof_on ? globaltimer_ns() : 0;
                    mbar_wait(&full[s], (parity >> s) & 1u);
                    if (prof_on) wait_c += globaltimer_ns() - tw;
                    parity ^= 1u << s;
                }
                const uint32_t sbase = ring_thread + (uint32_t)s * stage_bytes;
                CoarseStage co;
                co.gp = ring_u32 + (uint32_t)s * stage_bytes + off_cgp + (uint32_t)hsel * co_gp_bytes;
                co.idx = ring_u32 + (uint32_t)s * stage_bytes + off_cidx + (uint32_t)hsel * co_idx_bytes;
                co.base = 0;
                co.zslot = (FusedVar<VAR>::gp_staged && f.co_mode == 1) ? ((chunk / f.co_ipp) % p.wd) * (p.wh * p.ww) : 0;
#pragma unroll
                for (int k = 0; k < 4; k++) co.tab[k] = co_tab[k];
                if (nv == kItemVox) fused_item_apply<VAR, true, SILU>(p, f, sbase, off_g0, off_g1, c, ga, m1, m2, rk, hsel, t128, n, cqp, Cqp, total, v0, nv, dy, co);
                else fused_item_apply<VAR, false, SILU>(p, f, sbase, off_g0, off_g1, c, ga, m1, m2, rk, hsel, t128, n, cqp, Cqp, total, v0, nv, dy, co);
                if (j - S >= 0 || round + 1 < nrounds) {                      // the producer refills this stage
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                }
                if (--chunk < 0) { chunk = nchunks - 1; sp--; }
            }
        }
        if (prof_on) { prof[4] = globaltimer_ns(); prof[6] = wait_c; }
        named_bar_sync(kFusedBar, 256);               // gm / ctab / bound_s are rewritten by the next round
    }
    if (blockIdx.x != 0) return;
    // ---------------- CTA 0: the tensor's scale, and the parameter gradients from the complete sums (after the last grid
    // barrier every (n, c) entry is final): dgamma = sum dr xhat, dbeta = sum dr, dbias = sum dy
    if (threadIdx.x == 0) {
        if (dscale == 0.f) dscale = 1.f;
        p.dy_scale[0] = bound_max; p.dy_scale[1] = dscale; p.dy_scale[2] = 1.f / dscale;
    }
    if (!f.dgamma && !f.dbeta && !f.dbias) return;
    for (int c = threadIdx.x; c < p.C; c += 256) {
        const double ga = (p.gamma && f.mode != 0) ? (double)p.gamma[c] : 1.0;
        double dg = 0.0, db = 0.0, dbi = 0.0;
        const int cg = f.mode == 1 ? p.C / f.G : 1;
#pragma unroll 4
        for (int n = 0; n < p.N; n++) {
            const size_t i = (size_t)n * Cp + c;
            const double S1 = __ldcg(p.sums + i * 2), S2 = __ldcg(p.sums + i * 2 + 1);
            dg += S2; db += S1;
            if (f.dbias) {
                double m1 = 0.0, m2 = 0.0;
                if (f.mode == 2) { m1 = (double)gm[c].x; m2 = (double)gm[c].y; }         // (single round: still in place)
                else if (f.mode == 1 && gm_keep) { const float2 m = gm[n * f.G + c / cg]; m1 = (double)m.x; m2 = (double)m.y; }
                else if (f.mode == 1) {
                    // more (sample, group) pairs than the table holds: recompute from the sums
                    const int c_first = (c / cg) * cg;
                    for (int q = 0; q < cg; q++) {
                        const double gq = p.gamma ? (double)p.gamma[c_first + q] : 1.0;
                        m1 += gq * __ldcg(p.sums + ((size_t)n * Cp + c_first + q) * 2);
                        m2 += gq * __ldcg(p.sums + ((size_t)n * Cp + c_first + q) * 2 + 1);
                    }
                    m1 *= f.inv_count; m2 *= f.inv_count;
                }
                const double rr = p.rstd ? (double)p.rstd[i] : 1.0, mu = p.mean ? (double)p.mean[i] : 0.0;
                double sum_xhat = 0.0;
                if (f.mode == 1 || f.mode == 2) sum_xhat = rr * (f.fwd_stats[((size_t)n * p.C + c) * 2] - f.S * mu);
                dbi += rr * (ga * S1 - f.S * m1 - m2 * sum_xhat);
            }
        }
        if (f.dgamma) f.dgamma[c] = (float)dg;
        if (f.dbeta) f.dbeta[c] = (float)db;
        if (f.dbias) f.dbias[c] = (float)dbi;
    }
}

// ------------------------------------------------------------------------------------------------
// 1x1x1 head (+ softmax / argmax, crop-and-place)
// ------------------------------------------------------------------------------------------------
static constexpr int kHeadMaxCo = 16;

// CO: compile-time bound on the output channels (4 for the usual 2..4 classes: no predicated-off FMA slots)
template <int CO>
__global__ void __launch_bounds__(256) head_kernel(const e3b_head_args p, int Cq)
{
    extern __shared__ float sw[];      // [Co][Cq*4] weights then [Co] bias
    const int Cp = Cq * 4;
    for (int i = threadIdx.x; i < p.Co * Cp; i += blockDim.x) {
        const int co = i / Cp, c = i % Cp;
        sw[i] = c < p.C ? p.w[co * p.C + c] : 0.f;
    }
    for (int i = threadIdx.x; i < p.Co; i += blockDim.x) sw[p.Co * Cp + i] = p.b ? p.b[i] : 0.f;
    __syncthreads();
    const uint4* a = reinterpret_cast<const uint4*>(p.a);      // QH operand tensor: 16-byte units of 8 channels
    const int Ch = cpad16(p.C) / 8;
    const unsigned box = (unsigned)p.cn_d * p.cn_h * p.cn_w;   // (the launch wrapper checks that N * box fits 32 bits)
    const unsigned total = (unsigned)p.N * box;
    const size_t S = (size_t)p.D * p.H * p.W;
    const size_t Sd = (size_t)p.Dd * p.Hd * p.Wd;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = (int)(i / box);
        unsigned r = i - (unsigned)n * box;
        const int z = (int)(r / ((unsigned)p.cn_h * p.cn_w)); r -= (unsigned)z * p.cn_h * p.cn_w;
        const int y = (int)(r / (unsigned)p.cn_w);
        const int x = (int)(r - (unsigned)y * p.cn_w);
        // (flip: the features are the network output on a mirrored tile -- FlipAugment.backward of the TTA)
        const int zi = (p.flip & 1) ? p.D - 1 - (z + p.c0_d) : z + p.c0_d, yi = (p.flip & 2) ? p.H - 1 - (y + p.c0_h) : y + p.c0_h,
                  xi = (p.flip & 4) ? p.W - 1 - (x + p.c0_w) : x + p.c0_w;
        const size_t vin = ((size_t)zi * p.H + yi) * p.W + xi;
        float acc[CO];
#pragma unroll
        for (int co = 0; co < CO; co++) acc[co] = co < p.Co ? sw[p.Co * Cp + co] : 0.f;
        for (int cq = 0; cq < Cq; cq += 2) {
            // one full unit (8 channels) per load; Cq is even (channels are padded to 8)
            const uint4 u = a[((size_t)n * Ch + (cq >> 1)) * S + vin];
            const float4 v = unpack_half4(make_uint2(u.x, u.y)), v2 = unpack_half4(make_uint2(u.z, u.w));
#pragma unroll
            for (int co = 0; co < CO; co++) {
                if (co < p.Co) {
                    const float* wr = sw + co * Cp + cq * 4;
                    acc[co] = fmaf(v.x, wr[0], fmaf(v.y, wr[1], fmaf(v.z, wr[2], fmaf(v.w, wr[3], acc[co]))));
                    acc[co] = fmaf(v2.x, wr[4], fmaf(v2.y, wr[5], fmaf(v2.z, wr[6], fmaf(v2.w, wr[7], acc[co]))));
                }
            }
        }
        int oz = 0, oy = 0, ox = 0;
        if (p.dst_origin) { oz = p.dst_origin[3 * n]; oy = p.dst_origin[3 * n + 1]; ox = p.dst_origin[3 * n + 2]; }
        // tiles of a volume that is not a multiple of the tile shape overhang the destination
        // (the reference pads and slices back, inference.py:645-687,634): clip
        if (z + oz >= p.Dd || y + oy >= p.Hd || x + ox >= p.Wd || z + oz < 0 || y + oy < 0 || x + ox < 0) continue;
        const size_t vout = ((size_t)(z + oz) * p.Hd + (y + oy)) * p.Wd + (x + ox);
        const size_t nb = p.dst_single ? 0 : (size_t)n;
        if (p.out_mode == 1 || (p.out_mode == 2 && p.use_threshold)) {
            float mx = acc[0];
#pragma unroll
            for (int co = 1; co < CO; co++) if (co < p.Co) mx = fmaxf(mx, acc[co]);
            float sum = 0.f;
#pragma unroll
            for (int co = 0; co < CO; co++) if (co < p.Co) { acc[co] = expf(acc[co] - mx); sum += acc[co]; }
            const float inv = 1.f / sum;
#pragma unroll
            for (int co = 0; co < CO; co++) acc[co] *= inv;
        }
        if (p.out_mode == 2) {
            if (p.use_threshold) {
#pragma unroll
                for (int co = 0; co < CO; co++) if (co < p.Co && !(acc[co] > p.threshold)) acc[co] = 0.f;
            }
            int best = 0; float bv = acc[0];
#pragma unroll
            for (int co = 1; co < CO; co++) if (co < p.Co && acc[co] > bv) { bv = acc[co]; best = co; }
            reinterpret_cast<uint8_t*>(p.dst)[nb * Sd + vout] = (uint8_t)best;
        } else {
            float* d = reinterpret_cast<float*>(p.dst);
            const float scl = p.acc_scale != 0.f ? p.acc_scale : 1.f;
#pragma unroll
            for (int co = 0; co < CO; co++)
                if (co < p.Co) {
                    float v = acc[co];
                    if (p.round_half) v = __half2float(__float2half_rn(v));
                    float* o = d + (nb * p.Co + co) * Sd + vout;
                    *o = p.accumulate ? fmaf(scl, v, *o) : scl * v;
                }
        }
    }
}

// argmax over the channels of a probability volume, optional threshold (the Predictor's deferred argmax after the
// test-time-augmentation mean, inference.py:519-523)
__global__ void __launch_bounds__(256) prob_argmax_kernel(const float* __restrict__ prob, uint8_t* __restrict__ dst, int N, int C,
                                                          size_t S, int use_threshold, float threshold)
{
    const size_t total = (size_t)N * S;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t n = i / S, v = i % S;
        int best = 0; float bv = 0.f;
        for (int c = 0; c < C; c++) {
            float x = prob[(n * C + c) * S + v];
            if (use_threshold && !(x > threshold)) x = 0.f;
            if (c == 0 || x > bv) { bv = x; best = c; }
        }
        dst[i] = (uint8_t)best;
    }
}

// da[v][c] = sum_co dl[v][co] * w[co][c]
__global__ void __launch_bounds__(256) head_bwd_data_kernel(const float* __restrict__ dl, const float* __restrict__ w,
                                                            float4* __restrict__ da, int N, int C, int Cq, int Co, size_t S)
{
    extern __shared__ float sw[];
    const int Cp = Cq * 4;
    for (int i = threadIdx.x; i < Co * Cp; i += blockDim.x) {
        const int co = i / Cp, c = i % Cp;
        sw[i] = c < C ? w[co * C + c] : 0.f;
    }
    __syncthreads();
    const size_t total = (size_t)N * S;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t n = i / S, v = i % S;
        float g[kHeadMaxCo];
#pragma unroll
        for (int co = 0; co < kHeadMaxCo; co++) g[co] = co < Co ? dl[(n * Co + co) * S + v] : 0.f;
        for (int cq = 0; cq < Cq; cq++) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int co = 0; co < kHeadMaxCo; co++) {
                if (co < Co) {
                    const float* wr = sw + co * Cp + cq * 4;
                    o.x = fmaf(g[co], wr[0], o.x); o.y = fmaf(g[co], wr[1], o.y);
                    o.z = fmaf(g[co], wr[2], o.z); o.w = fmaf(g[co], wr[3], o.w);
                }
            }
            da[(n * Cq + cq) * S + v] = o;
        }
    }
}

// ws[co][c] += sum_v dl[v][co]*a[v][c] ; ws[Co*Cp + co] += sum_v dl[v][co].   grid (chunks, Cq)
// One pass over the activations: every thread keeps all Co partial sums of its 4 channels.
template <int CO>
__global__ void __launch_bounds__(256) head_bwd_w_kernel(const float* __restrict__ dl, const uint2* __restrict__ a,
                                                         double* __restrict__ ws, int N, int Cq, int Ch, int Co, size_t S)
{
    const int cq = blockIdx.y, Cp = Cq * 4;
    const size_t total = (size_t)N * S;
    float acc[CO][4], sb[CO];
#pragma unroll
    for (int co = 0; co < CO; co++) { acc[co][0] = acc[co][1] = acc[co][2] = acc[co][3] = 0.f; sb[co] = 0.f; }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t n = i / S, v = i % S;
        const float4 av = unpack_half4(a[qh_index((int)n, Ch, cq, S, v)]);
#pragma unroll
        for (int co = 0; co < CO; co++) {
            if (co < Co) {
                const float g = dl[(n * Co + co) * S + v];
                acc[co][0] = fmaf(g, av.x, acc[co][0]); acc[co][1] = fmaf(g, av.y, acc[co][1]);
                acc[co][2] = fmaf(g, av.z, acc[co][2]); acc[co][3] = fmaf(g, av.w, acc[co][3]);
                sb[co] += g;
            }
        }
    }
    __shared__ float red[8][CO * 5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int co = 0; co < CO; co++) {
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int j = 0; j < 4; j++) acc[co][j] += __shfl_xor_sync(0xffffffffu, acc[co][j], o);
            sb[co] += __shfl_xor_sync(0xffffffffu, sb[co], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 4; j++) red[warp][co * 5 + j] = acc[co][j];
            red[warp][co * 5 + 4] = sb[co];
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < Co * 5; t += blockDim.x) {
        double sum = 0.0;
        for (int w = 0; w < 8; w++) sum += (double)red[w][t];
        const int co = t / 5, j = t % 5;
        if (j < 4) atomicAdd(ws + (size_t)co * Cp + cq * 4 + j, sum);
        else if (cq == 0) atomicAdd(ws + (size_t)Co * Cp + co, sum);
    }
}

// <= 4 classes: one thread per 16-byte unit (8 channels of one voxel): full-sector loads of the activations, the
// logit gradients are re-read once per 8 channels instead of once per 4, 32-bit indices.  grid: (chunks, Ch, N)
__global__ void __launch_bounds__(256) head_bwd_w8_kernel(const float* __restrict__ dl, const uint4* __restrict__ a,
                                                          double* __restrict__ ws, int Cp, int Ch, int Co, int S)
{
    constexpr int CO = 4;
    const int ch = blockIdx.y, n = blockIdx.z;
    float acc[CO][8], sb[CO];
#pragma unroll
    for (int co = 0; co < CO; co++) {
        sb[co] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) acc[co][j] = 0.f;
    }
    const uint4* ap = a + ((size_t)n * Ch + ch) * S;
    const float* gp = dl + (size_t)n * Co * S;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < S; v += gridDim.x * blockDim.x) {
        const uint4 u = ap[v];
        const float4 lo = unpack_half4(make_uint2(u.x, u.y)), hi = unpack_half4(make_uint2(u.z, u.w));
#pragma unroll
        for (int co = 0; co < CO; co++) {
            if (co < Co) {
                const float g = gp[(size_t)co * S + v];
                acc[co][0] = fmaf(g, lo.x, acc[co][0]); acc[co][1] = fmaf(g, lo.y, acc[co][1]);
                acc[co][2] = fmaf(g, lo.z, acc[co][2]); acc[co][3] = fmaf(g, lo.w, acc[co][3]);
                acc[co][4] = fmaf(g, hi.x, acc[co][4]); acc[co][5] = fmaf(g, hi.y, acc[co][5]);
                acc[co][6] = fmaf(g, hi.z, acc[co][6]); acc[co][7] = fmaf(g, hi.w, acc[co][7]);
                sb[co] += g;
            }
        }
    }
    __shared__ float red[8][CO * 9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int co = 0; co < CO; co++) {
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int j = 0; j < 8; j++) acc[co][j] += __shfl_xor_sync(0xffffffffu, acc[co][j], o);
            sb[co] += __shfl_xor_sync(0xffffffffu, sb[co], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 8; j++) red[warp][co * 9 + j] = acc[co][j];
            red[warp][co * 9 + 8] = sb[co];
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < Co * 9; t += blockDim.x) {
        double sum = 0.0;
        for (int w = 0; w < 8; w++) sum += (double)red[w][t];
        const int co = t / 9, j = t % 9;
        if (j < 8) { if (ch * 8 + j < Cp) atomicAdd(ws + (size_t)co * Cp + ch * 8 + j, sum); }
        else if (ch == 0) atomicAdd(ws + (size_t)Co * Cp + co, sum);
    }
}

__global__ void head_bwd_finish_kernel(const double* __restrict__ ws, float* __restrict__ dw, float* __restrict__ db, int C,
                                       int Cp, int Co)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Co * C) dw[i] = (float)ws[(size_t)(i / C) * Cp + (i % C)];
    if (i < Co) db[i] = (float)ws[(size_t)Co * Cp + i];
}

}  // namespace e3b

// =================================================================================================
// C ABI
// =================================================================================================
using namespace e3b;

extern "C" {

int e3b_pack_ncdhw(const float* src, void* dst_qp, int N, int C, int D, int H, int W, int Dv, int Hv,
                   int Wv, int z0, int y0, int x0, void* stream)
{
    if (N <= 0 || C <= 0) return set_error("pack: empty tensor");
    const int Ch = cpad16(C) / 8;
    if ((long long)D * H * W >= (1ll << 31) || Ch > 65535 || N > 65535) return set_error("pack: tensor too large for the launch grid");
    const dim3 grid(grid_for((size_t)D * H * W, 256, 4), Ch, N);
    pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, nullptr, reinterpret_cast<uint4*>(dst_qp), C, Ch, D, H, W, Dv, Hv, Wv,
                                                        z0, y0, x0, 0, 0);
    return check_launch("pack_ncdhw");
}

int e3b_gather_tiles(const float* vol, const int32_t* origins, void* dst_qp, int B, int C, int D, int H, int W, int Dv,
                     int Hv, int Wv, int flip, void* stream)
{
    if (B <= 0 || C <= 0) return set_error("gather: empty batch");
    const int Ch = cpad16(C) / 8;
    if ((long long)D * H * W >= (1ll << 31) || Ch > 65535 || B > 65535) return set_error("gather: tile too large for the launch grid");
    const dim3 grid(grid_for((size_t)D * H * W, 256, 4), Ch, B);
    pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(vol, origins, reinterpret_cast<uint4*>(dst_qp), C, Ch, D, H, W, Dv, Hv, Wv,
                                                        0, 0, 0, 1, flip);
    return check_launch("gather_tiles");
}

int e3b_unpack_qp(const float* src_qp, float* dst, int N, int C, int D, int H, int W, void* stream)
{
    const size_t S = (size_t)D * H * W;
    unpack_kernel<<<grid_for((size_t)N * C * S, 256), 256, 0, (cudaStream_t)stream>>>(src_qp, dst, N, C, cpad8(C) / 4, S);
    return check_launch("unpack_qp");
}

static bool pack_mode_ok(int mode, const PackDims& d, int kd, int kh, int kw)
{
    if (mode < 0 || mode > 5 || d.NT <= 0) return false;
    if (mode >= 4 && (kd != 3 || kh != 3 || kw != 3)) return false;
    return true;
}

int64_t e3b_packed_weight_floats(int mode, int C0, int C1, int Co, int kd, int kh, int kw)
{
    if (mode < 0 || mode > 5) return -1;
    PackDims d = pack_dims(mode, C0, C1, Co, kd, kh, kw);
    if (!pack_mode_ok(mode, d, kd, kh, kw)) return -1;
    return (int64_t)d.ktot * d.ntot * d.taps / 2;        // fp16 elements, counted in floats
}

int e3b_pack_weights(int mode, const float* w, const float* scale, const float* wscale, void* dst, int C0, int C1, int Co,
                     int kd, int kh, int kw, void* stream)
{
    if (mode < 0 || mode > 5) return set_error("pack_weights: bad mode %d", mode);
    PackDims d = pack_dims(mode, C0, C1, Co, kd, kh, kw);
    if (!pack_mode_ok(mode, d, kd, kh, kw)) return set_error("pack_weights: mode %d does not support output width %d / these taps", mode, d.ntot);
    const size_t total = (size_t)d.ktot * d.ntot * d.taps;
    pack_weights_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(mode, w, scale, wscale, reinterpret_cast<__half*>(dst), C0,
                                                                                C1, Co, kd * kh * kw, d);
    return check_launch("pack_weights");
}

int e3b_pack_jobs_fill(const e3b_pack_job* jobs, int njobs, void* host_table, int64_t* total_blocks)
{
    if (!jobs || njobs <= 0 || !host_table || !total_blocks) return set_error("pack_jobs_fill: bad arguments");
    PackJobDev* out = reinterpret_cast<PackJobDev*>(host_table);
    int64_t blocks = 0;
    for (int i = 0; i < njobs; i++) {
        const e3b_pack_job& j = jobs[i];
        if (j.mode < 0 || j.mode > 5) return set_error("pack_jobs_fill: bad mode %d", j.mode);
        PackDims d = pack_dims(j.mode, j.C0, j.C1, j.Co, j.kd, j.kh, j.kw);
        if (!pack_mode_ok(j.mode, d, j.kd, j.kh, j.kw)) return set_error("pack_jobs_fill: job %d: unsupported configuration", i);
        PackJobDev& o = out[i];
        o.w = j.w; o.scale = j.scale; o.wscale = j.wscale; o.dst = reinterpret_cast<__half*>(j.dst);
        o.mode = j.mode; o.C0 = j.C0; o.C1 = j.C1; o.Co = j.Co; o.tu = j.kd * j.kh * j.kw; o.d = d;
        o.total = (unsigned long long)d.ktot * d.ntot * d.taps;
        o.first_block = (int)blocks;
        blocks += (int64_t)((o.total + 255) / 256);
        if (blocks > 0x7fffffff) return set_error("pack_jobs_fill: too many elements");
    }
    *total_blocks = blocks;
    return 0;
}

int64_t e3b_pack_job_table_bytes(int njobs) { return (int64_t)njobs * (int64_t)sizeof(PackJobDev); }

int e3b_pack_weights_batched(const void* device_table, int njobs, int64_t total_blocks, void* stream)
{
    if (!device_table || njobs <= 0 || total_blocks <= 0) return set_error("pack_weights_batched: bad arguments");
    pack_weights_batched_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const PackJobDev*>(device_table), njobs);
    return check_launch("pack_weights_batched");
}

int e3b_weight_scales(const e3b_ws_job* device_jobs, int njobs, float* table, uint32_t* scratch, void* stream)
{
    if (!device_jobs || !table || !scratch || njobs <= 0 || njobs > 65535) return set_error("weight_scales: bad arguments");
    weight_scales_kernel<<<dim3(kWsChunks, njobs), 256, 0, (cudaStream_t)stream>>>(device_jobs, table, scratch);
    return check_launch("weight_scales");
}

int e3b_norm_finalize(const double* stats, int mode, int G, int N, int C, int64_t S, const float* gamma, const float* beta,
                      float eps, float* running_mean, float* running_var, float momentum, float* scale, float* shift,
                      float* mean, float* rstd, void* stream)
{
    if (mode == 1 && (G <= 0 || C % G)) return set_error("norm: num_channels %d not divisible by num_groups %d", C, G);
    if (mode == 3 && (!running_mean || !running_var)) return set_error("norm: eval-mode batch norm needs running stats");
    const int Cp = cpad8(C);
    norm_finalize_kernel<<<(N * Cp + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        stats, mode, G, N, C, Cp, (double)S, gamma, beta, eps, running_mean, running_var, momentum, scale, shift, mean, rstd);
    return check_launch("norm_finalize");
}

int e3b_norm_act(const void* y, const float* scale, const float* shift, void* a, void* pooled, uint8_t* pool_idx, int N, int C,
                 int D, int H, int W, int pk_d, int pk_h, int pk_w, int relu, float act_slope, const float* act_slope_dev, int y_is_half,
                 void* stream)
{
    if (relu < 0 || relu > 2) return set_error("norm_act: unknown activation code %d", relu);
    if (act_slope_dev && relu != 1) return set_error("norm_act: a device-resident slope belongs to activation code 1");
    const bool pooling = pooled || pool_idx;
    if (y_is_half && (!pooling || scale || a)) return set_error("norm_act: an fp16 input is only pooled (eval path)");
    if (!pooling) { pk_d = pk_h = pk_w = 1; }
    if (pk_d < 1 || pk_h < 1 || pk_w < 1 || pk_d > 2 || pk_h > 2 || pk_w > 2) return set_error("norm_act: pooling kernel must be 1 or 2 per dim");
    if ((scale == nullptr) != (shift == nullptr)) return set_error("norm_act: scale and shift go together");
    NormActDev p;
    p.y = y_is_half ? nullptr : reinterpret_cast<const float4*>(y); p.scale = scale; p.shift = shift;
    p.yh = y_is_half ? reinterpret_cast<const uint2*>(y) : nullptr;
    p.a = reinterpret_cast<uint2*>(a); p.pooled = reinterpret_cast<uint2*>(pooled);
    p.pool_idx = reinterpret_cast<uchar4*>(pool_idx);
    p.C = C; p.N = N; p.Cq = cpad8(C) / 4; p.Ch = cpad16(C) / 8; p.D = D; p.H = H; p.W = W; p.pkd = pk_d; p.pkh = pk_h; p.pkw = pk_w; p.relu = relu; p.slope = act_slope; p.slope_dev = act_slope_dev;
    p.Dp = (D + pk_d - 1) / pk_d; p.Hp = (H + pk_h - 1) / pk_h; p.Wp = (W + pk_w - 1) / pk_w;
    if (p.Cq > 65535 || N > 65535) return set_error("norm_act: too many channels / samples for the launch grid");
    if (pooling && y_is_half && !pool_idx && pooled && !getenv("E3B_POOL_GENERIC")) {
        // inference: pooling only, whole 16-byte units
        const size_t Sp = (size_t)p.Dp * p.Hp * p.Wp;
        size_t chunks = (Sp + 255) / 256; if (chunks > 8192) chunks = 8192;
        pool_qh_kernel<<<dim3((unsigned)chunks, p.Ch, N), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const uint4*>(y), reinterpret_cast<uint4*>(pooled), p.Ch, D, H, W, pk_d, pk_h, pk_w, p.Dp, p.Hp, p.Wp);
    } else if (pooling) {
        const dim3 grid((unsigned)(p.Dp * ((p.Hp * p.Wp + 255) / 256)), p.Cq, N);
        norm_act_pool_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    } else {
        const size_t S = (size_t)D * H * W;
        const dim3 grid((unsigned)((S + 256 * kVpt - 1) / (256 * kVpt)), p.Cq, N);
        norm_act_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    }
    return check_launch("norm_act");
}

static int fill_bwd(const e3b_norm_bwd_args* a, NormBwdDev& p)
{
    p.y = reinterpret_cast<const float4*>(a->y);
    p.scale = a->scale; p.shift = a->shift;
    if ((a->scale == nullptr) != (a->shift == nullptr)) return set_error("norm_bwd: scale and shift go together");
    p.g0 = reinterpret_cast<const float4*>(a->g0); p.g1 = reinterpret_cast<const float4*>(a->g1);
    p.gp = reinterpret_cast<const float4*>(a->gp);
    p.pool_idx = reinterpret_cast<const uchar4*>(a->pool_idx);
    if (a->gp && !a->pool_idx) return set_error("norm_bwd: a pooled gradient needs the forward pool_idx");
    p.N = a->N; p.C = a->C; p.Cq = cpad8(a->C) / 4; p.D = a->D; p.H = a->H; p.W = a->W;
    p.wd = p.wh = p.ww = 1;
    if (a->gp && a->s2d) return set_error("norm_bwd: pooled gradient and space-to-depth output are exclusive");
    if (a->gp) { p.wd = a->pk_d; p.wh = a->pk_h; p.ww = a->pk_w; }
    if (a->s2d) { p.wd = a->sd; p.wh = a->sh; p.ww = a->sw; }
    if (p.wd < 1 || p.wh < 1 || p.ww < 1 || p.wd * p.wh * p.ww > 8) return set_error("norm_bwd: window > 8 voxels");
    p.Dw = (a->D + p.wd - 1) / p.wd; p.Hw = (a->H + p.wh - 1) / p.wh; p.Ww = (a->W + p.ww - 1) / p.ww;
    p.Dg = a->D; p.Hg = a->H; p.Wg = a->W;
    if (a->s2d) { p.Dg = p.Dw * p.wd; p.Hg = p.Hw * p.wh; p.Wg = p.Ww * p.ww; }
    if (a->relu < 0 || a->relu > 2) return set_error("norm_bwd: unknown activation code %d", a->relu);
    p.relu = a->relu; p.slope = a->act_slope; p.s2d = a->s2d;
    p.slope_dev = a->act_slope_dev; p.slope_sums = a->slope_sums;
    if (a->act_slope_dev && a->relu != 1) return set_error("norm_bwd: a device-resident slope belongs to activation code 1");
    if (a->dslope && !a->slope_sums) return set_error("norm_bwd: dslope needs the slope_sums workspace");
    p.gamma = (a->mode == 0) ? nullptr : a->gamma;
    p.mean = a->mean; p.rstd = a->rstd; p.m1 = a->m1; p.m2 = a->m2;
    p.sums = a->sums; p.dy = reinterpret_cast<uint2*>(a->dy);
    p.amax = a->amax; p.dy_scale = a->dy_scale;
    p.Ch = a->s2d ? cpad16(p.wd * p.wh * p.ww * p.Cq * 4) / 8 : cpad16(a->C) / 8;
    p.g1_crop = a->g1_crop; p.g1_od = a->g1_od; p.g1_oh = a->g1_oh; p.g1_ow = a->g1_ow;
    p.g1_D = a->g1_D; p.g1_H = a->g1_H; p.g1_W = a->g1_W;
    if (a->g1_crop && (!a->g1 || a->g1_od < 0 || a->g1_oh < 0 || a->g1_ow < 0 || a->g1_od + a->g1_D > a->D ||
                       a->g1_oh + a->g1_H > a->H || a->g1_ow + a->g1_W > a->W))
        return set_error("norm_bwd: the cropped skip gradient box does not fit the tensor");
    if (p.Cq > 65535 || a->N > 65535) return set_error("norm_bwd: too many channels / samples for the launch grid");
    return 0;
}

int e3b_norm_bwd_reduce(const e3b_norm_bwd_args* a, void* stream)
{
    NormBwdDev p;
    if (fill_bwd(a, p)) return 1;
    const int Cp = p.Cq * 4;
    if (!a->amax || !a->dy_scale) return set_error("norm_bwd: amax / dy_scale buffers are required");
    const size_t sums_b = sizeof(double) * 2 * (size_t)a->N * Cp, amax_b = sizeof(unsigned int) * 2 * (size_t)a->N * Cp;
    cudaError_t e;
    if ((char*)a->amax == (char*)a->sums + sums_b && (char*)a->dy_scale == (char*)a->amax + amax_b) {
        // the three accumulators sit back to back (engine.py allocates them so): one memset node instead of three
        e = cudaMemsetAsync(a->sums, 0, sums_b + amax_b + sizeof(float) * 4, (cudaStream_t)stream);
    } else {
        e = cudaMemsetAsync(a->sums, 0, sums_b, (cudaStream_t)stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(a->amax, 0, amax_b, (cudaStream_t)stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(a->dy_scale, 0, sizeof(float) * 4, (cudaStream_t)stream);
    }
    if (e == cudaSuccess && a->slope_sums) e = cudaMemsetAsync(a->slope_sums, 0, sizeof(double) * (size_t)a->N * Cp, (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error("memset: %s", cudaGetErrorString(e));
    const size_t vox = (size_t)p.D * p.H * p.W;
    int bx = (int)((vox + 256 * kRedVpt - 1) / (256 * kRedVpt));
    int cap = (16 * num_sms()) / (p.Cq * a->N); if (cap < 1) cap = 1;
    if (bx > cap) bx = cap;
    norm_bwd_reduce_kernel<<<dim3(bx, p.Cq, a->N), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("norm_bwd_reduce");
}

int e3b_norm_bwd_finalize(const e3b_norm_bwd_args* a, void* stream)
{
    const int Cp = cpad8(a->C);
    if (a->mode == 1 && (a->G <= 0 || a->C % a->G)) return set_error("norm_bwd: bad group count");
    if (a->dbias && (a->mode == 1 || a->mode == 2) && !a->fwd_stats) return set_error("norm_bwd: dbias needs fwd_stats");
    const int nt = a->N * Cp;
    if (a->dslope) {
        if (!a->slope_sums) return set_error("norm_bwd: dslope needs the slope_sums workspace");
        prelu_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a->slope_sums, nt, a->dslope);
        if (check_launch("prelu_finalize")) return 1;
    }
    if (nt <= 1024) {
        const int threads = (nt + 31) & ~31;
        norm_bwd_finalize_block_kernel<<<1, threads, sizeof(double) * 3 * (size_t)nt, (cudaStream_t)stream>>>(
            a->sums, a->fwd_stats, a->mode, a->G, a->N, a->C, Cp, (double)a->D * a->H * a->W, a->gamma, a->mean, a->rstd,
            a->m1, a->m2, a->dgamma, a->dbeta, a->dbias, a->amax, a->dy_scale);
        return check_launch("norm_bwd_finalize");
    }
    norm_bwd_finalize_kernel<<<(nt + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        a->sums, a->fwd_stats, a->mode, a->G, a->N, a->C, Cp, (double)a->D * a->H * a->W, a->gamma, a->mean, a->rstd, a->m1,
        a->m2, a->dgamma, a->dbeta, a->dbias, a->amax, a->dy_scale);
    return check_launch("norm_bwd_finalize");
}

int e3b_norm_bwd_apply(const e3b_norm_bwd_args* a, void* stream)
{
    NormBwdDev p;
    if (fill_bwd(a, p)) return 1;
    // vector path: rows of 4-voxel groups
    if (!p.s2d && (p.W % 4 == 0)) {
        const size_t groups = (size_t)p.D * p.H * (p.W / 4);
        const dim3 grid((unsigned)((groups + 255) / 256), p.Cq, a->N);
        norm_bwd_apply_x4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
        return check_launch("norm_bwd_apply_x4");
    }
    const size_t Sg = (size_t)p.Dg * p.Hg * p.Wg;
    const dim3 grid((unsigned)((Sg + 256 * kAppVpt - 1) / (256 * kAppVpt)), p.Cq, a->N);
    norm_bwd_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("norm_bwd_apply");
}

int e3b_norm_bwd_fused(const e3b_norm_bwd_args* a, void* stream)
{
    NormBwdDev p;
    if (fill_bwd(a, p)) return 1;
    const int Cp = p.Cq * 4;
    if (!a->amax || !a->dy_scale || !a->sums) return set_error("norm_bwd_fused: sums / amax / dy_scale buffers are required");
    if (a->act_slope_dev || a->dslope) return set_error("norm_bwd_fused: nn.PReLU is served by the reduce / finalize / apply kernels");
    if (Cp > kFusedMaxCp) return set_error("norm_bwd_fused: more than %d channels: use the reduce / finalize / apply kernels", kFusedMaxCp);
    if (a->mode == 1 && (a->G <= 0 || a->C % a->G)) return set_error("norm_bwd: bad group count");
    if (a->dbias && (a->mode == 1 || a->mode == 2) && !a->fwd_stats) return set_error("norm_bwd: dbias needs fwd_stats");
    if (a->s2d && (a->D % a->sd || a->H % a->sh || a->W % a->sw))
        return set_error("norm_bwd_fused: space-to-depth output needs extents divisible by the stride (use the three-kernel path)");
    const size_t sums_b = sizeof(double) * 2 * (size_t)a->N * Cp, amax_b = sizeof(unsigned int) * 2 * (size_t)a->N * Cp;
    if (!((char*)a->amax == (char*)a->sums + sums_b && (char*)a->dy_scale == (char*)a->amax + amax_b))
        return set_error("norm_bwd_fused: sums, amax and dy_scale must be one contiguous workspace");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(a->sums, 0, sums_b + amax_b + sizeof(float) * 4, st);
    if (e != cudaSuccess) return set_error("memset: %s", cudaGetErrorString(e));
    FusedDev f;
    f.mode = a->mode; f.G = a->G; f.S = (double)a->D * a->H * a->W; f.fwd_stats = a->fwd_stats;
    f.dgamma = a->dgamma; f.dbeta = a->dbeta; f.dbias = a->dbias;
    f.counter = reinterpret_cast<unsigned int*>(a->dy_scale + 3);
    f.per_sample = a->mode != 2;
    f.Cp = Cp;
    f.prof = getenv("E3B_FUSED_PROF") != nullptr;
    if (f.prof) {
        static const unsigned long long zeros[64] = {0};
        cudaMemcpyToSymbolAsync(g_fused_prof, zeros, sizeof(zeros), 0, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    }
    f.inv_count = a->mode == 1 ? 1.0 / (f.S * (a->C / a->G)) : (a->mode == 2 ? 1.0 / (f.S * a->N) : 1.0);
    f.dHW = make_fastdiv(p.H * p.W); f.dW = make_fastdiv(p.W);
    f.dwd = make_fastdiv(p.wd); f.dwh = make_fastdiv(p.wh); f.dww = make_fastdiv(p.ww);
    // persistent grid of co-resident CTAs (the kernel contains grid barriers): a cooperative launch, which the driver
    // only starts when every CTA fits at once -- also next to kernels of other streams
    f.g1_bulk = (p.g1 != nullptr && !p.g1_crop) ? 1 : 0;
    int var = 4;
    if (p.g0 && !p.gp && !p.g1_crop && !p.s2d) var = f.g1_bulk ? 1 : 0;
    else if (p.g0 && !p.g1 && !p.gp && p.s2d) var = 2;
    else if (!p.g0 && f.g1_bulk && p.gp && !p.s2d) var = 3;
    f.co_n = 0; f.co_ipp = 1; f.co_mode = 0;
    if (var == 3 && !getenv("E3B_FUSED_NO_STAGED_POOL")) {
        // can an item (512 consecutive voxels) be made of whole pooling windows whose coarse voxels are contiguous?
        const int HW = p.H * p.W;
        const bool even = p.D % p.wd == 0 && p.H % p.wh == 0 && p.W % p.ww == 0 && p.W % 8 == 0;
        if (even && HW % kItemVox == 0 && kItemVox % p.W == 0 && (kItemVox / p.W) % p.wh == 0 && p.wd <= 2) {
            f.co_mode = 1; f.co_ipp = HW / kItemVox; f.co_n = (kItemVox / p.W / p.wh) * p.Ww;
        } else if (even && kItemVox % HW == 0 && (kItemVox / HW) % p.wd == 0 && p.D % (kItemVox / HW) == 0) {
            f.co_mode = 2; f.co_n = (kItemVox / HW / p.wd) * p.Hw * p.Ww;
        }
        if (f.co_mode && (f.co_n % 4 || f.co_n > 128 || f.co_n < 4)) f.co_mode = 0;
        if (f.co_mode) var = 5;
    }
    const int nb = 1 + (p.g0 != nullptr) + f.g1_bulk;
    const int kRingBytes = 96 * 1024;               // two CTAs per SM
    const size_t stage_bytes = (size_t)nb * kItemTensorBytes + (var == 5 ? (size_t)2 * f.co_n * 20 : 0);
    f.stages = (int)(kRingBytes / stage_bytes);
    if (f.stages > kFusedMaxStages) f.stages = kFusedMaxStages;
    const size_t smem = (size_t)f.stages * stage_bytes;
    typedef void (*FusedKernel)(const NormBwdDev, const FusedDev);
    static const FusedKernel kernels[2][6] = {
        {norm_bwd_fused_kernel<0, false>, norm_bwd_fused_kernel<1, false>, norm_bwd_fused_kernel<2, false>,
         norm_bwd_fused_kernel<3, false>, norm_bwd_fused_kernel<4, false>, norm_bwd_fused_kernel<5, false>},
        {norm_bwd_fused_kernel<0, true>, norm_bwd_fused_kernel<1, true>, norm_bwd_fused_kernel<2, true>,
         norm_bwd_fused_kernel<3, true>, norm_bwd_fused_kernel<4, true>, norm_bwd_fused_kernel<5, true>}};
    const int silu = a->relu == 2 ? 1 : 0;
    const FusedKernel kern = kernels[silu][var];
    var += 6 * silu;                                // (index of the per-kernel occupancy cache)
    static int per_sm[kMaxDevices][12][4] = {};
    const int dev = current_device();
    if (!per_sm[dev][var][nb]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes) != cudaSuccess)
            return set_error("norm_bwd_fused: cannot reserve %d bytes of shared memory", kRingBytes);
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kFusedThreads, smem) != cudaSuccess || n < 1)
            return set_error("norm_bwd_fused: occupancy query failed");
        per_sm[dev][var][nb] = n > 2 ? 2 : n;
    }
    // samples per round: as many as keep the round's working set (sources + dy) inside L2, a divisor of N
    f.nset = a->N;
    if (f.per_sample) {
        const double per_sample = (double)Cp * f.S * (4.0 * nb + 2.0);
        const double budget = 100e6;                // of the 126 MB L2
        long long lim = (long long)(budget / per_sample);
        if (lim > kCtabSlots / p.Cq) lim = kCtabSlots / p.Cq;
        if (lim < 1) lim = 1;
        if (lim > a->N) lim = a->N;
        while (a->N % lim) lim--;
        f.nset = (int)lim;
    }
    const long long items = (long long)f.nset * (p.Cq / 2) * (((long long)p.D * p.H * p.W + kItemVox - 1) / kItemVox);
    if (items >= (1ll << 31)) return set_error("norm_bwd_fused: tensor too large");
    long long grid = (long long)num_sms() * per_sm[dev][var][nb];
    if (grid > items) grid = items;
    if (grid < 1) grid = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kFusedThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, p, f);
    if (e != cudaSuccess) return set_error("norm_bwd_fused launch: %s", cudaGetErrorString(e));
    return check_launch("norm_bwd_fused");
}

int e3b_debug_fused_prof(unsigned long long* out64)
{
    cudaError_t e = cudaMemcpyFromSymbol(out64, g_fused_prof, sizeof(unsigned long long) * 64);
    if (e != cudaSuccess) return set_error("e3b_debug_fused_prof: %s", cudaGetErrorString(e));
    return 0;
}

int e3b_head(const e3b_head_args* a, void* stream)
{
    if (a->Co > kHeadMaxCo) return set_error("head: out_channels %d > %d not supported", a->Co, kHeadMaxCo);
    if (a->out_mode < 0 || a->out_mode > 2) return set_error("head: bad out_mode");
    if (a->out_mode == 2 && (a->accumulate || a->round_half)) return set_error("head: accumulate / round_half apply to float outputs");
    const int Cq = cpad8(a->C) / 4;
    const size_t total = (size_t)a->N * a->cn_d * a->cn_h * a->cn_w;
    if (total >= ((size_t)1 << 32)) return set_error("head: more than 2^32 output voxels in one call");
    const size_t smem = sizeof(float) * ((size_t)a->Co * Cq * 4 + a->Co);
    if (a->Co <= 4) head_kernel<4><<<grid_for(total, 256), 256, smem, (cudaStream_t)stream>>>(*a, Cq);
    else head_kernel<kHeadMaxCo><<<grid_for(total, 256), 256, smem, (cudaStream_t)stream>>>(*a, Cq);
    return check_launch("head");
}

int e3b_prob_argmax(const float* prob, uint8_t* dst, int N, int C, int64_t S, int use_threshold, float threshold, void* stream)
{
    if (N <= 0 || C <= 0 || C > 255 || S <= 0) return set_error("prob_argmax: bad extents");
    prob_argmax_kernel<<<grid_for((size_t)N * S, 256), 256, 0, (cudaStream_t)stream>>>(prob, dst, N, C, (size_t)S, use_threshold, threshold);
    return check_launch("prob_argmax");
}

int e3b_head_bwd(const float* dl, const void* a, const float* w, float* da, float* dw, float* db, double* workspace, int N,
                 int C, int Co, int D, int H, int W, void* stream)
{
    if (Co > kHeadMaxCo) return set_error("head_bwd: out_channels %d > %d not supported", Co, kHeadMaxCo);
    const int Cq = cpad8(C) / 4, Cp = Cq * 4;
    const size_t S = (size_t)D * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    if (da) {
        head_bwd_data_kernel<<<grid_for((size_t)N * S, 256), 256, sizeof(float) * Co * Cp, st>>>(
            dl, w, reinterpret_cast<float4*>(da), N, C, Cq, Co, S);
        if (check_launch("head_bwd_data")) return 1;
    }
    cudaError_t e = cudaMemsetAsync(workspace, 0, sizeof(double) * ((size_t)Co * Cp + Co), st);
    if (e != cudaSuccess) return set_error("memset: %s", cudaGetErrorString(e));
    int bx = (8 * num_sms()) / Cq; if (bx < 1) bx = 1;
    const int Ch = cpad16(C) / 8;
    if (Co <= 4 && S < ((size_t)1 << 31) && N <= 65535) {
        int b8 = (8 * num_sms()) / (Ch * N); if (b8 < 1) b8 = 1;
        head_bwd_w8_kernel<<<dim3(b8, Ch, N), 256, 0, st>>>(dl, reinterpret_cast<const uint4*>(a), workspace, Cp, Ch, Co, (int)S);
    } else if (Co <= 4) head_bwd_w_kernel<4><<<dim3(bx, Cq), 256, 0, st>>>(dl, reinterpret_cast<const uint2*>(a), workspace, N, Cq, Ch, Co, S);
    else head_bwd_w_kernel<kHeadMaxCo><<<dim3(bx, Cq), 256, 0, st>>>(dl, reinterpret_cast<const uint2*>(a), workspace, N, Cq, Ch, Co, S);
    if (check_launch("head_bwd_w")) return 1;
    head_bwd_finish_kernel<<<(Co * C + 127) / 128 + 1, 128, 0, st>>>(workspace, dw, db, C, Cp, Co);
    return check_launch("head_bwd_finish");
}

}  // extern "C"
