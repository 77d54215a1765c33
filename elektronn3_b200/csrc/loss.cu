// Generalised Dice loss of the reference (elektronn3 modules/loss.py:165-233) fused into two passes over the logits:
//   forward : softmax over the channels + one-hot target + the three per-class sums (intersection, prob, target)
//             in ONE read of the logits (the reference materialises probs, the one-hot tensor, their product and
//             their sum: ~10 launches over (N, C, spatial) tensors), then loss = mean_c w_c (1 - (2 I_c + s) / (P_c + T_c + s + eps));
//   backward: d loss / d logits in ONE read + one write (softmax recomputed, no saved probabilities).
// HBM-bound: 4 C + 8 bytes read per voxel forward, 4 C + 8 read + 4 C written backward.
#include "common.cuh"
#include "kernels.h"

namespace e3b {

static constexpr int kLossMaxC = 16;

template <bool SOFTMAX>
E3B_DEVINL void load_probs(const float* __restrict__ x, size_t base, size_t S, int C, float* p)
{
    float mx = -3.4e38f;
#pragma unroll
    for (int c = 0; c < kLossMaxC; c++) if (c < C) { p[c] = x[base + (size_t)c * S]; mx = fmaxf(mx, p[c]); }
    if (SOFTMAX) {
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < kLossMaxC; c++) if (c < C) { p[c] = expf(p[c] - mx); sum += p[c]; }
        const float inv = 1.f / sum;
#pragma unroll
        for (int c = 0; c < kLossMaxC; c++) if (c < C) p[c] *= inv;
    }
}

static constexpr int kVecC = 4;            // the vector path serves <= 4 classes
// four consecutive voxels of one sample: softmax probabilities p[k][c] from float4 loads per channel
template <bool SOFTMAX>
E3B_DEVINL void load_probs4(const float* __restrict__ x, size_t base, size_t S, int C, float (*p)[kVecC])
{
    float mx[4] = {-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f};
#pragma unroll
    for (int c = 0; c < kVecC; c++) {
        if (c < C) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + base + (size_t)c * S));
            p[0][c] = v.x; p[1][c] = v.y; p[2][c] = v.z; p[3][c] = v.w;
            mx[0] = fmaxf(mx[0], v.x); mx[1] = fmaxf(mx[1], v.y); mx[2] = fmaxf(mx[2], v.z); mx[3] = fmaxf(mx[3], v.w);
        }
    }
    if (SOFTMAX) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < kVecC; c++) if (c < C) { p[k][c] = expf(p[k][c] - mx[k]); sum += p[k][c]; }
            const float inv = 1.f / sum;
#pragma unroll
            for (int c = 0; c < kVecC; c++) if (c < C) p[k][c] *= inv;
        }
    }
}

E3B_DEVINL void load_targets4(const long long* __restrict__ target, size_t i, int* t)
{
    const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(target + i)), b = __ldg(reinterpret_cast<const longlong2*>(target + i + 2));
    t[0] = (int)a.x; t[1] = (int)a.y; t[2] = (int)b.x; t[3] = (int)b.y;
}

// sums[3][C] (fp64 atomics): I_c = sum p_c t_c, P_c = sum p_c, T_c = sum t_c.  target: dense int64 (N, S) or one-hot float (N, C, S)
// grid: (chunks, N); VEC: four voxels per thread and iteration (S % 4 == 0, dense targets)
template <bool SOFTMAX, bool VEC>
__global__ void __launch_bounds__(256) dice_fwd_kernel(const float* __restrict__ x, const long long* __restrict__ target,
                                                       const float* __restrict__ onehot, int N, int C, size_t S, double* __restrict__ sums)
{
    float aI[kLossMaxC], aP[kLossMaxC], aT[kLossMaxC];
#pragma unroll
    for (int c = 0; c < kLossMaxC; c++) { aI[c] = 0.f; aP[c] = 0.f; aT[c] = 0.f; }
    const size_t n = blockIdx.y;
    if (VEC) {
        for (unsigned v = (blockIdx.x * blockDim.x + threadIdx.x) * 4u; v < (unsigned)S; v += gridDim.x * blockDim.x * 4u) {
            float p[4][kVecC];
            int t[4];
            load_targets4(target, n * S + v, t);
            load_probs4<SOFTMAX>(x, n * C * S + v, S, C, p);
#pragma unroll
            for (int k = 0; k < 4; k++) {
#pragma unroll
                for (int c = 0; c < kVecC; c++) {
                    if (c < C) {
                        const float tc = c == t[k] ? 1.f : 0.f;
                        aI[c] = fmaf(p[k][c], tc, aI[c]); aP[c] += p[k][c]; aT[c] += tc;
                    }
                }
            }
        }
    } else {
        for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < (unsigned)S; v += gridDim.x * blockDim.x) {
            const size_t base = n * C * S + v;
            float p[kLossMaxC];
            load_probs<SOFTMAX>(x, base, S, C, p);
            const int t = target ? (int)target[n * S + v] : -1;
#pragma unroll
            for (int c = 0; c < kLossMaxC; c++) {
                if (c < C) {
                    const float tc = target ? (c == t ? 1.f : 0.f) : onehot[base + (size_t)c * S];
                    aI[c] = fmaf(p[c], tc, aI[c]); aP[c] += p[c]; aT[c] += tc;
                }
            }
        }
    }
    __shared__ float red[8][3 * kLossMaxC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < kLossMaxC; c++) {
        if (c < C) {
            for (int o = 16; o > 0; o >>= 1) {
                aI[c] += __shfl_xor_sync(0xffffffffu, aI[c], o);
                aP[c] += __shfl_xor_sync(0xffffffffu, aP[c], o);
                aT[c] += __shfl_xor_sync(0xffffffffu, aT[c], o);
            }
            if (lane == 0) { red[warp][c] = aI[c]; red[warp][kLossMaxC + c] = aP[c]; red[warp][2 * kLossMaxC + c] = aT[c]; }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 3 * C; t += blockDim.x) {
        const int which = t / C, c = t - which * C;
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += (double)red[w][which * kLossMaxC + c];
        atomicAdd(sums + which * C + c, s);
    }
}

// loss and the per-class coefficients of d loss / d p_c(v) = a_c t_c(v) + b_c
__global__ void dice_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ weight, int weight_n, int C,
                                     double smooth, double eps, float* __restrict__ loss, float* __restrict__ coef)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double acc = 0.0;
    for (int c = 0; c < C; c++) {
        const double w = weight ? (double)weight[weight_n > 1 ? c : 0] : 1.0;
        const double num = 2.0 * sums[c] + smooth, den = sums[C + c] + sums[2 * C + c] + smooth + eps;
        acc += w * (1.0 - num / den);
        coef[c] = (float)(-(w / C) * 2.0 / den);              // a_c
        coef[C + c] = (float)((w / C) * num / (den * den));   // b_c
    }
    loss[0] = (float)(acc / C);
}

template <bool SOFTMAX, bool VEC>
__global__ void __launch_bounds__(256) dice_bwd_kernel(const float* __restrict__ x, const long long* __restrict__ target,
                                                       const float* __restrict__ onehot, const float* __restrict__ coef,
                                                       const float* __restrict__ gout, float* __restrict__ dx, int N, int C, size_t S)
{
    __shared__ float sc[2 * kLossMaxC];
    if (threadIdx.x < 2 * C) sc[threadIdx.x] = coef[threadIdx.x] * gout[0];
    __syncthreads();
    const size_t n = blockIdx.y;
    if (VEC) {
        for (unsigned v = (blockIdx.x * blockDim.x + threadIdx.x) * 4u; v < (unsigned)S; v += gridDim.x * blockDim.x * 4u) {
            float p[4][kVecC];
            int t[4];
            const size_t base = n * C * S + v;
            load_targets4(target, n * S + v, t);
            load_probs4<SOFTMAX>(x, base, S, C, p);
            float dot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 4; k++) {
#pragma unroll
                for (int c = 0; c < kVecC; c++)
                    if (c < C) dot[k] = fmaf(p[k][c], fmaf(sc[c], c == t[k] ? 1.f : 0.f, sc[C + c]), dot[k]);
            }
#pragma unroll
            for (int c = 0; c < kVecC; c++) {
                if (c < C) {
                    float o[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float g = fmaf(sc[c], c == t[k] ? 1.f : 0.f, sc[C + c]);
                        o[k] = SOFTMAX ? p[k][c] * (g - dot[k]) : g;
                    }
                    *reinterpret_cast<float4*>(dx + base + (size_t)c * S) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    } else {
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < (unsigned)S; v += gridDim.x * blockDim.x) {
        const size_t base = n * C * S + v;
        float p[kLossMaxC], g[kLossMaxC];
        load_probs<SOFTMAX>(x, base, S, C, p);
        const int t = target ? (int)target[n * S + v] : -1;
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < kLossMaxC; c++) {
            if (c < C) {
                const float tc = target ? (c == t ? 1.f : 0.f) : onehot[base + (size_t)c * S];
                g[c] = fmaf(sc[c], tc, sc[C + c]);
                dot = fmaf(p[c], g[c], dot);
            }
        }
#pragma unroll
        for (int c = 0; c < kLossMaxC; c++)
            if (c < C) dx[base + (size_t)c * S] = SOFTMAX ? p[c] * (g[c] - dot) : g[c];
    }
    }
}

static dim3 loss_grid(int N, size_t S, int waves, int vec)
{
    size_t b = (S / vec + 255) / 256;
    size_t cap = ((size_t)num_sms() * waves + N - 1) / N;
    if (cap < 1) cap = 1;
    return dim3((unsigned)(b > cap ? cap : (b < 1 ? 1 : b)), (unsigned)N);
}

}  // namespace e3b

using namespace e3b;

extern "C" {

int e3b_dice_fwd(const float* logits, const int64_t* target, const float* target_onehot, const float* weight, int weight_n, int N,
                 int C, int64_t S, int apply_softmax, float smooth, float eps, double* sums, float* loss, float* coef, void* stream)
{
    if (!logits || (!target && !target_onehot) || !sums || !loss || !coef) return set_error("dice: null pointer");
    if (C < 1 || C > kLossMaxC) return set_error("dice: 1..%d classes supported, got %d", kLossMaxC, C);
    if (N <= 0 || S <= 0) return set_error("dice: empty tensor");
    if (N > 65535 || S >= ((int64_t)1 << 32)) return set_error("dice: tensor too large for the launch grid");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 3 * C, st);
    if (e != cudaSuccess) return set_error("memset: %s", cudaGetErrorString(e));
    const long long* t = reinterpret_cast<const long long*>(target);
    // vector path: four voxels per thread (dense targets, S % 4 == 0, <= 4 classes: register budget); two CTAs per SM, so
    // that only ~300 CTAs add their partial sums to the 3 C fp64 accumulators
    const bool vec = t != nullptr && S % 4 == 0 && C <= 4 && ((uintptr_t)logits % 16 == 0) && ((uintptr_t)t % 16 == 0);
    if (vec) {
        if (apply_softmax) dice_fwd_kernel<true, true><<<loss_grid(N, (size_t)S, 2, 4), 256, 0, st>>>(logits, t, target_onehot, N, C, (size_t)S, sums);
        else dice_fwd_kernel<false, true><<<loss_grid(N, (size_t)S, 2, 4), 256, 0, st>>>(logits, t, target_onehot, N, C, (size_t)S, sums);
    } else {
        if (apply_softmax) dice_fwd_kernel<true, false><<<loss_grid(N, (size_t)S, 4, 1), 256, 0, st>>>(logits, t, target_onehot, N, C, (size_t)S, sums);
        else dice_fwd_kernel<false, false><<<loss_grid(N, (size_t)S, 4, 1), 256, 0, st>>>(logits, t, target_onehot, N, C, (size_t)S, sums);
    }
    if (check_launch("dice_fwd")) return 1;
    dice_finalize_kernel<<<1, 32, 0, st>>>(sums, weight, weight_n, C, (double)smooth, (double)eps, loss, coef);
    return check_launch("dice_finalize");
}

int e3b_dice_bwd(const float* logits, const int64_t* target, const float* target_onehot, const float* coef, const float* gout,
                 float* dlogits, int N, int C, int64_t S, int apply_softmax, void* stream)
{
    if (!logits || (!target && !target_onehot) || !coef || !gout || !dlogits) return set_error("dice_bwd: null pointer");
    if (C < 1 || C > kLossMaxC) return set_error("dice_bwd: 1..%d classes supported, got %d", kLossMaxC, C);
    cudaStream_t st = (cudaStream_t)stream;
    if (N > 65535 || S >= ((int64_t)1 << 32)) return set_error("dice_bwd: tensor too large for the launch grid");
    const long long* t = reinterpret_cast<const long long*>(target);
    const bool vec = t != nullptr && S % 4 == 0 && C <= 4 && ((uintptr_t)logits % 16 == 0) && ((uintptr_t)t % 16 == 0) &&
                     ((uintptr_t)dlogits % 16 == 0);
    if (vec) {
        if (apply_softmax) dice_bwd_kernel<true, true><<<loss_grid(N, (size_t)S, 8, 4), 256, 0, st>>>(logits, t, target_onehot, coef, gout, dlogits, N, C, (size_t)S);
        else dice_bwd_kernel<false, true><<<loss_grid(N, (size_t)S, 8, 4), 256, 0, st>>>(logits, t, target_onehot, coef, gout, dlogits, N, C, (size_t)S);
    } else {
        if (apply_softmax) dice_bwd_kernel<true, false><<<loss_grid(N, (size_t)S, 8, 1), 256, 0, st>>>(logits, t, target_onehot, coef, gout, dlogits, N, C, (size_t)S);
        else dice_bwd_kernel<false, false><<<loss_grid(N, (size_t)S, 8, 1), 256, 0, st>>>(logits, t, target_onehot, coef, gout, dlogits, N, C, (size_t)S);
    }
    return check_launch("dice_bwd");
}

}  // extern "C"
