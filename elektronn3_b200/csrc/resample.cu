// Element-wise kernels of the UNet options off the default configuration:
//   * merge_mode='add'   (models/unet.py:399-401): updec + centre-cropped skip tensor, both QH operand tensors
//   * up_mode='resizeconv_*' (models/unet.py:411-449, ResizeConv): nn.Upsample(scale_factor=2 or (1,2,2), 'nearest' |
//     'trilinear' / 'bilinear', align_corners=False) in front of a conv3 / conv1.  The up-sampled QH tensor is written with
//     the conv's zero padding made explicit (and autocrop's trailing-voxel crop of the conv OUTPUT folded in), so that the
//     convolution that follows is a plain VALID one on the existing tensor-core kernels; the backward gathers the
//     gradient of that padded tensor back onto the coarse grid (the transpose of the interpolation).
//   * the residual shortcut of resunet's ConvBlock (models/resunet.py:252-261, `y += proj(inp)` between conv2 and norm2):
//     y += r with the statistics of the SUM for the norm that follows (e3b_residual_add), and the accumulation of the
//     shortcut's gradient into the gradient of the block input (e3b_qp_axpy).
// All of them are HBM bound: one thread per 16-byte unit, coalesced along x.
#include "common.cuh"
#include "kernels.h"

namespace e3b {

E3B_DEVINL void unpack8(const uint4& u, float* f)
{
    const float4 lo = unpack_half4(make_uint2(u.x, u.y)), hi = unpack_half4(make_uint2(u.z, u.w));
    f[0] = lo.x; f[1] = lo.y; f[2] = lo.z; f[3] = lo.w; f[4] = hi.x; f[5] = hi.y; f[6] = hi.z; f[7] = hi.w;
}
E3B_DEVINL uint4 pack8(const float* f)
{
    const uint2 lo = pack_half4(f[0], f[1], f[2], f[3]), hi = pack_half4(f[4], f[5], f[6], f[7]);
    return make_uint4(lo.x, lo.y, hi.x, hi.y);
}

// dst = a + b[centre crop]; a, dst (N, Ch, D, H, W) units, b (N, Ch, D1, H1, W1) read at the voxel offset (od, oh, ow).
// grid: (voxel chunks, Ch, N)
__global__ void __launch_bounds__(256) qh_add_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ dst,
                                                     int Ch, int D, int H, int W, int D1, int H1, int W1, int od, int oh, int ow)
{
    const int S = D * H * W, HW = H * W;
    const int ch = blockIdx.y, n = blockIdx.z;
    const size_t pa = ((size_t)n * Ch + ch) * (size_t)S;
    const size_t pb = ((size_t)n * Ch + ch) * ((size_t)D1 * H1 * W1);
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < S; v += gridDim.x * blockDim.x) {
        const int z = v / HW, r = v - z * HW, y = r / W, x = r - y * W;
        const uint4 ua = __ldg(a + pa + v);
        const uint4 ub = __ldg(b + pb + ((size_t)(z + od) * H1 + (y + oh)) * W1 + (x + ow));
        float fa[8], fb[8];
        unpack8(ua, fa); unpack8(ub, fb);
#pragma unroll
        for (int j = 0; j < 8; j++) fa[j] += fb[j];
        dst[pa + v] = pack8(fa);
    }
}

// One axis of nn.Upsample(scale_factor=s, align_corners=False): fine index f -> the two coarse taps (i0, i1) and their
// weights (w0, w1).  s = 1 is the identity.  Nearest: i0 = f / s.  Linear: src = (f + 0.5) / s - 0.5 clamped at 0
// (ATen area_pixel_compute_source_index), i1 = min(i0 + 1, n - 1).
struct AxisTap { int i0, i1; float w0, w1; };
E3B_DEVINL AxisTap axis_tap(int f, int s, int n, int linear)
{
    AxisTap t;
    if (s == 1) { t.i0 = t.i1 = f; t.w0 = 1.f; t.w1 = 0.f; return t; }
    if (!linear) { t.i0 = t.i1 = min(f / s, n - 1); t.w0 = 1.f; t.w1 = 0.f; return t; }
    float src = ((float)f + 0.5f) / (float)s - 0.5f;
    if (src < 0.f) src = 0.f;
    t.i0 = min((int)src, n - 1);
    t.i1 = min(t.i0 + 1, n - 1);
    t.w1 = src - (float)t.i0; t.w0 = 1.f - t.w1;
    return t;
}

struct UpDev {
    int Ch, d, h, w;                 // coarse extents (units of 16 bytes: Ch planes of 8 channels, or Cq planes of 4 floats)
    int Dp, Hp, Wp;                  // extents of the padded fine tensor
    int sd, sh, sw;                  // scale per axis (1 or 2)
    int od, oh, ow;                  // where fine voxel 0 sits in the padded tensor (the conv's zero padding: 0 or 1)
    int Rd, Rh, Rw;                  // fine voxels [0, R) per axis carry data; everything else of the padded tensor is 0
    int linear;
};

// coarse QH (N, Ch, d, h, w) -> padded fine QH (N, Ch, Dp, Hp, Wp);  grid: (voxel chunks of the fine tensor, Ch, N)
__global__ void __launch_bounds__(256) qh_upsample_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const UpDev p)
{
    const int S = p.Dp * p.Hp * p.Wp, HW = p.Hp * p.Wp;
    const int ch = blockIdx.y, n = blockIdx.z;
    const uint4* sb = src + ((size_t)n * p.Ch + ch) * ((size_t)p.d * p.h * p.w);
    uint4* db = dst + ((size_t)n * p.Ch + ch) * (size_t)S;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < S; v += gridDim.x * blockDim.x) {
        const int zp = v / HW, r = v - zp * HW, yp = r / p.Wp, xp = r - yp * p.Wp;
        const int fz = zp - p.od, fy = yp - p.oh, fx = xp - p.ow;
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (fz >= 0 && fz < p.Rd && fy >= 0 && fy < p.Rh && fx >= 0 && fx < p.Rw) {
            const AxisTap tz = axis_tap(fz, p.sd, p.d, p.linear), ty = axis_tap(fy, p.sh, p.h, p.linear),
                          tx = axis_tap(fx, p.sw, p.w, p.linear);
            if (!p.linear) {
                out = __ldg(sb + ((size_t)tz.i0 * p.h + ty.i0) * p.w + tx.i0);
            } else {
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const float wgt = ((c & 4) ? tz.w1 : tz.w0) * ((c & 2) ? ty.w1 : ty.w0) * ((c & 1) ? tx.w1 : tx.w0);
                    if (wgt == 0.f) continue;
                    const int iz = (c & 4) ? tz.i1 : tz.i0, iy = (c & 2) ? ty.i1 : ty.i0, ix = (c & 1) ? tx.i1 : tx.i0;
                    float f[8];
                    unpack8(__ldg(sb + ((size_t)iz * p.h + iy) * p.w + ix), f);
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[j] = fmaf(wgt, f[j], acc[j]);
                }
                out = pack8(acc);
            }
        }
        db[v] = out;
    }
}

// gradient of the padded fine tensor, QP fp32 (N, Cq, Dp, Hp, Wp) -> gradient of the coarse tensor, QP (N, Cq, d, h, w):
// every coarse voxel gathers the fine voxels it contributed to (the transpose of qh_upsample_kernel).
__global__ void __launch_bounds__(256) qp_upsample_bwd_kernel(const float4* __restrict__ gfine, float4* __restrict__ gcoarse, const UpDev p)
{
    const int S = p.d * p.h * p.w, HW = p.h * p.w;
    const int cq = blockIdx.y, n = blockIdx.z;
    const float4* gb = gfine + ((size_t)n * p.Ch + cq) * ((size_t)p.Dp * p.Hp * p.Wp);
    float4* ob = gcoarse + ((size_t)n * p.Ch + cq) * (size_t)S;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < S; v += gridDim.x * blockDim.x) {
        const int z = v / HW, r = v - z * HW, y = r / p.w, x = r - y * p.w;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        // fine voxels that can read coarse voxel c along an axis of scale s: [s c - (s - 1), s c + 2 s - 2] (linear), [s c, s c + s - 1] (nearest)
        const int ez = p.linear ? p.sd - 1 : 0, ey = p.linear ? p.sh - 1 : 0, ex = p.linear ? p.sw - 1 : 0;
        for (int fz = max(0, z * p.sd - ez); fz <= min(p.Rd - 1, z * p.sd + p.sd - 1 + ez); fz++) {
            const AxisTap tz = axis_tap(fz, p.sd, p.d, p.linear);
            const float wz = (tz.i0 == z ? tz.w0 : 0.f) + (tz.i1 == z ? tz.w1 : 0.f);
            if (wz == 0.f) continue;
            for (int fy = max(0, y * p.sh - ey); fy <= min(p.Rh - 1, y * p.sh + p.sh - 1 + ey); fy++) {
                const AxisTap ty = axis_tap(fy, p.sh, p.h, p.linear);
                const float wy = (ty.i0 == y ? ty.w0 : 0.f) + (ty.i1 == y ? ty.w1 : 0.f);
                if (wy == 0.f) continue;
                for (int fx = max(0, x * p.sw - ex); fx <= min(p.Rw - 1, x * p.sw + p.sw - 1 + ex); fx++) {
                    const AxisTap tx = axis_tap(fx, p.sw, p.w, p.linear);
                    const float wx = (tx.i0 == x ? tx.w0 : 0.f) + (tx.i1 == x ? tx.w1 : 0.f);
                    if (wx == 0.f) continue;
                    const float wgt = wz * wy * wx;
                    const float4 g = __ldg(gb + ((size_t)(fz + p.od) * p.Hp + (fy + p.oh)) * p.Wp + (fx + p.ow));
                    acc.x = fmaf(wgt, g.x, acc.x); acc.y = fmaf(wgt, g.y, acc.y);
                    acc.z = fmaf(wgt, g.z, acc.z); acc.w = fmaf(wgt, g.w, acc.w);
                }
            }
        }
        ob[v] = acc;
    }
}

static int fill_up(UpDev& p, int planes, int d, int h, int w, int Dp, int Hp, int Wp, int sd, int sh, int sw, int od, int oh, int ow,
                   int Rd, int Rh, int Rw, int linear, const char* what)
{
    if (d <= 0 || h <= 0 || w <= 0 || Dp <= 0 || Hp <= 0 || Wp <= 0) return set_error("%s: empty tensor", what);
    if (sd < 1 || sd > 2 || sh < 1 || sh > 2 || sw < 1 || sw > 2) return set_error("%s: scale factors must be 1 or 2", what);
    if (od < 0 || oh < 0 || ow < 0) return set_error("%s: negative offset", what);
    if (Rd < 0 || Rd > d * sd || Rh < 0 || Rh > h * sh || Rw < 0 || Rw > w * sw) return set_error("%s: data range exceeds the up-sampled extents", what);
    if (od + Rd > Dp || oh + Rh > Hp || ow + Rw > Wp) return set_error("%s: data range does not fit the padded tensor", what);
    p.Ch = planes; p.d = d; p.h = h; p.w = w; p.Dp = Dp; p.Hp = Hp; p.Wp = Wp; p.sd = sd; p.sh = sh; p.sw = sw;
    p.od = od; p.oh = oh; p.ow = ow; p.Rd = Rd; p.Rh = Rh; p.Rw = Rw; p.linear = linear ? 1 : 0;
    return 0;
}

// y (QP fp32, in place) += r, r a QP fp32 tensor or a QH fp16 operand tensor of the same extents (N, C, S voxels);
// stats (optional): per (n, c) sum and sum of squares of the result (fp64 atomics; zeroed by the caller).
// grid: (voxel chunks, Cq, N)
__global__ void __launch_bounds__(256) residual_add_kernel(float4* __restrict__ y, const float4* __restrict__ r32,
                                                           const uint2* __restrict__ r16, double* __restrict__ stats, int C, int Cq,
                                                           int Ch, int S)
{
    const int cq = blockIdx.y, n = blockIdx.z;
    float4* yb = y + ((size_t)n * Cq + cq) * (size_t)S;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < S; v += gridDim.x * blockDim.x) {
        float4 a = yb[v];
        float4 b;
        if (r32) b = __ldg(r32 + ((size_t)n * Cq + cq) * (size_t)S + v);
        else b = unpack_half4(__ldg(r16 + (((size_t)n * Ch + (cq >> 1)) * (size_t)S + v) * 2 + (cq & 1)));
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        yb[v] = a;
        s1[0] += a.x; s1[1] += a.y; s1[2] += a.z; s1[3] += a.w;
        s2[0] = fmaf(a.x, a.x, s2[0]); s2[1] = fmaf(a.y, a.y, s2[1]); s2[2] = fmaf(a.z, a.z, s2[2]); s2[3] = fmaf(a.w, a.w, s2[3]);
    }
    if (!stats) return;
    __shared__ float red[8][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        for (int o = 16; o > 0; o >>= 1) {
            s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
            s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 4; j++) { red[warp][j] = s1[j]; red[warp][4 + j] = s2[j]; }
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += (double)red[w][threadIdx.x];
        const int c = cq * 4 + (threadIdx.x & 3);
        if (c < C) atomicAdd(stats + ((size_t)n * C + c) * 2 + (threadIdx.x >> 2), t);
    }
}

// dst (QP fp32, in place) += alpha * src; src QP fp32 or QH fp16; alpha a device scalar (nullptr: 1)
__global__ void __launch_bounds__(256) qp_axpy_kernel(float4* __restrict__ dst, const float4* __restrict__ s32, const uint2* __restrict__ s16,
                                                      const float* __restrict__ alpha, int Cq, int Ch, int S)
{
    const int cq = blockIdx.y, n = blockIdx.z;
    const float al = alpha ? __ldg(alpha) : 1.f;
    float4* db = dst + ((size_t)n * Cq + cq) * (size_t)S;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < S; v += gridDim.x * blockDim.x) {
        float4 a = db[v];
        float4 b;
        if (s32) b = __ldg(s32 + ((size_t)n * Cq + cq) * (size_t)S + v);
        else b = unpack_half4(__ldg(s16 + (((size_t)n * Ch + (cq >> 1)) * (size_t)S + v) * 2 + (cq & 1)));
        a.x = fmaf(al, b.x, a.x); a.y = fmaf(al, b.y, a.y); a.z = fmaf(al, b.z, a.z); a.w = fmaf(al, b.w, a.w);
        db[v] = a;
    }
}

static unsigned chunks_for(size_t S) { size_t c = (S + 255) / 256; return (unsigned)(c > 4096 ? 4096 : c); }

}  // namespace e3b

using namespace e3b;

extern "C" {

int e3b_add_qh(const void* a, const void* b, void* dst, int N, int C, int D, int H, int W, int D1, int H1, int W1,
               int off_d, int off_h, int off_w, void* stream)
{
    if (!a || !b || !dst) return set_error("add_qh: null tensor pointer");
    if (N <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return set_error("add_qh: empty tensor");
    if (off_d < 0 || off_h < 0 || off_w < 0 || off_d + D > D1 || off_h + H > H1 || off_w + W > W1)
        return set_error("add_qh: the cropped box does not fit the second tensor");
    const int Ch = cpad16(C) / 8;
    if (Ch > 65535 || N > 65535) return set_error("add_qh: too many channels / samples for the launch grid");
    qh_add_kernel<<<dim3(chunks_for((size_t)D * H * W), Ch, N), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), reinterpret_cast<uint4*>(dst), Ch, D, H, W, D1, H1, W1,
        off_d, off_h, off_w);
    return check_launch("add_qh");
}

int e3b_upsample_qh(const void* src, void* dst, int N, int C, int d, int h, int w, int Dp, int Hp, int Wp, int sd, int sh, int sw,
                    int off_d, int off_h, int off_w, int Rd, int Rh, int Rw, int linear, void* stream)
{
    if (!src || !dst) return set_error("upsample_qh: null tensor pointer");
    if (N <= 0 || C <= 0) return set_error("upsample_qh: empty tensor");
    UpDev p;
    const int Ch = cpad16(C) / 8;
    if (fill_up(p, Ch, d, h, w, Dp, Hp, Wp, sd, sh, sw, off_d, off_h, off_w, Rd, Rh, Rw, linear, "upsample_qh")) return 1;
    if (Ch > 65535 || N > 65535) return set_error("upsample_qh: too many channels / samples for the launch grid");
    qh_upsample_kernel<<<dim3(chunks_for((size_t)Dp * Hp * Wp), Ch, N), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), p);
    return check_launch("upsample_qh");
}

int e3b_upsample_bwd_qp(const float* gfine, float* gcoarse, int N, int C, int d, int h, int w, int Dp, int Hp, int Wp, int sd, int sh,
                        int sw, int off_d, int off_h, int off_w, int Rd, int Rh, int Rw, int linear, void* stream)
{
    if (!gfine || !gcoarse) return set_error("upsample_bwd_qp: null tensor pointer");
    if (N <= 0 || C <= 0) return set_error("upsample_bwd_qp: empty tensor");
    UpDev p;
    const int Cq = cpad8(C) / 4;
    if (fill_up(p, Cq, d, h, w, Dp, Hp, Wp, sd, sh, sw, off_d, off_h, off_w, Rd, Rh, Rw, linear, "upsample_bwd_qp")) return 1;
    if (Cq > 65535 || N > 65535) return set_error("upsample_bwd_qp: too many channels / samples for the launch grid");
    qp_upsample_bwd_kernel<<<dim3(chunks_for((size_t)d * h * w), Cq, N), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(gfine), reinterpret_cast<float4*>(gcoarse), p);
    return check_launch("upsample_bwd_qp");
}

int e3b_residual_add(float* y_qp, const void* r, int r_is_half, double* stats, int N, int C, int64_t S, void* stream)
{
    if (!y_qp || !r) return set_error("residual_add: null tensor pointer");
    if (N <= 0 || C <= 0 || S <= 0 || S >= (1ll << 31)) return set_error("residual_add: bad extents");
    const int Cq = cpad8(C) / 4, Ch = cpad16(C) / 8;
    if (Cq > 65535 || N > 65535) return set_error("residual_add: too many channels / samples for the launch grid");
    cudaStream_t st = (cudaStream_t)stream;
    if (stats) {
        cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)N * C, st);
        if (e != cudaSuccess) return set_error("memset: %s", cudaGetErrorString(e));
    }
    // (few chunks per slab: every block ends in 8 fp64 atomics per statistic)
    size_t chunks = ((size_t)S + 256 * 8 - 1) / (256 * 8);
    const size_t cap = (size_t)(16 * num_sms()) / ((size_t)Cq * N) + 1;
    if (chunks > cap) chunks = cap;
    residual_add_kernel<<<dim3((unsigned)chunks, Cq, N), 256, 0, st>>>(
        reinterpret_cast<float4*>(y_qp), r_is_half ? nullptr : reinterpret_cast<const float4*>(r),
        r_is_half ? reinterpret_cast<const uint2*>(r) : nullptr, stats, C, Cq, Ch, (int)S);
    return check_launch("residual_add");
}

int e3b_qp_axpy(float* dst_qp, const void* src, int src_is_half, const float* alpha, int N, int C, int64_t S, void* stream)
{
    if (!dst_qp || !src) return set_error("qp_axpy: null tensor pointer");
    if (N <= 0 || C <= 0 || S <= 0 || S >= (1ll << 31)) return set_error("qp_axpy: bad extents");
    const int Cq = cpad8(C) / 4, Ch = cpad16(C) / 8;
    if (Cq > 65535 || N > 65535) return set_error("qp_axpy: too many channels / samples for the launch grid");
    qp_axpy_kernel<<<dim3(chunks_for((size_t)S), Cq, N), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4*>(dst_qp), src_is_half ? nullptr : reinterpret_cast<const float4*>(src),
        src_is_half ? reinterpret_cast<const uint2*>(src) : nullptr, alpha, Cq, Ch, (int)S);
    return check_launch("qp_axpy");
}

}  // extern "C"
