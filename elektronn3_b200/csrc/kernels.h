// Internal declarations shared by the translation units of libe3b.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/e3b.h"

namespace e3b {

// thread-local error message; returns a non-zero code so callers can `return set_error(...)`
int set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);

static constexpr int kMaxDevices = 64;
int current_device();     // cudaGetDevice(), clamped to [0, kMaxDevices)
int num_sms();            // of the current device
int conv_ntile_width(int npad_total);
int make_qp_tensor_map(CUtensorMap* map, const float* ptr, int N, int Cq, int D, int H, int W, int Da, int Ha, int Wa,
                       int bx, int by, int bz, int bcq);
CUtensorMapL2promotion tma_l2_promotion();
int conv_debug_read(unsigned long long* out16, int reset);
int launch_conv_tc(const e3b_conv_args* a, cudaStream_t stream);
int conv_zs_supported(int C0, int C1, int n_total, int kd, int kh, int kw, int scatter);
int conv_zs_ntile(int C0, int C1, int n_total);
int conv_zs_debug_read(uint32_t* out, int n);
int conv_zs_prof_read(unsigned long long* out16, int reset);
int launch_conv_zs(const e3b_conv_args* a, cudaStream_t stream);
int launch_wgrad_tc(const e3b_wgrad_args* a, cudaStream_t stream);
int launch_wgrad_reduce_batched(const e3b_wgrad_args* args, int n, cudaStream_t stream);
int64_t wgrad_workspace_floats(const e3b_wgrad_args* a);

__host__ __device__ static inline int cpad8(int c) { return (c + 7) & ~7; }
__host__ __device__ static inline int cpad16(int c) { return (c + 15) & ~15; }

}  // namespace e3b
