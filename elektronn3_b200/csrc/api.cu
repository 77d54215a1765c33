// C-ABI plumbing of libe3b.so: error reporting, launch accounting, argument validation for the
// tensor-core entry points.  See include/e3b.h for the contract.
#include <atomic>
#include <stdarg.h>
#include <stdio.h>
#include "kernels.h"

namespace e3b {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    count_launch(1);
    return 0;
}

}  // namespace e3b

using namespace e3b;

extern "C" {

int e3b_version(void) { return E3B_VERSION; }
const char* e3b_last_error(void) { return g_err; }
int64_t e3b_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int e3b_conv(const e3b_conv_args* a, void* stream)
{
    if (!a) return set_error("conv: null args");
    if (!a->src0 || !a->dst0 || !a->wpk) return set_error("conv: null tensor pointer");
    if (a->N <= 0 || a->D <= 0 || a->H <= 0 || a->W <= 0 || a->C0 <= 0) return set_error("conv: empty input");
    if (a->pd < 0 || a->pd > 2 || a->ph < 0 || a->ph > 2 || a->pw < 0 || a->pw > 2) return set_error("conv: padding must be 0..2");
    if (a->src1 && (a->D1 < a->D || a->H1 < a->H || a->W1 < a->W)) return set_error("conv: source 1 smaller than source 0");
    if (a->scatter) {
        if (a->kd * a->kh * a->kw != 1) return set_error("conv: scatter mode is a 1-tap GEMM");
        if (a->sd * a->sh * a->sw > 8 || a->sd < 1 || a->sh < 1 || a->sw < 1) return set_error("conv: bad scatter strides");
        if (a->n_total != a->sd * a->sh * a->sw * cpad16(a->Cd0)) return set_error("conv: scatter n_total mismatch");
        if (a->dst1) return set_error("conv: scatter mode has one output");
    }
    if (((uintptr_t)a->src0 | (uintptr_t)a->dst0 | (uintptr_t)a->wpk | (uintptr_t)a->src1 | (uintptr_t)a->dst1) & 15)
        return set_error("conv: pointers must be 16-byte aligned");
    if (a->variant == 1) {
        if (!conv_zs_supported(a->C0, a->src1 ? a->C1 : 0, a->n_total, a->kd, a->kh, a->kw, a->scatter))
            return set_error("conv: variant 1 (z-stacked) does not support this configuration");
        return launch_conv_zs(a, (cudaStream_t)stream);
    }
    if (a->variant != 0) return set_error("conv: unknown variant %d", a->variant);
    return launch_conv_tc(a, (cudaStream_t)stream);
}

int e3b_conv_variant(int C0, int C1, int n_total, int kd, int kh, int kw, int scatter)
{
    return conv_zs_supported(C0, C1, n_total, kd, kh, kw, scatter);
}

int e3b_debug_zs_read(uint32_t* out, int n) { return conv_zs_debug_read(out, n); }
int e3b_debug_zs_prof(unsigned long long* out16, int reset) { return conv_zs_prof_read(out16, reset); }
int e3b_debug_conv_counters(unsigned long long* out16, int reset) { return conv_debug_read(out16, reset); }

int64_t e3b_wgrad_workspace_floats(const e3b_wgrad_args* a)
{
    if (!a) return -1;
    return wgrad_workspace_floats(a);
}

int e3b_wgrad(const e3b_wgrad_args* a, void* stream)
{
    if (!a) return set_error("wgrad: null args");
    if (!a->src0 || !a->dy || !a->dw || !a->workspace) return set_error("wgrad: null tensor pointer");
    if (a->N <= 0 || a->D <= 0 || a->H <= 0 || a->W <= 0 || a->C0 <= 0 || a->Co <= 0) return set_error("wgrad: empty input");
    if (!((a->kd == 1 || a->kd == 3) && (a->kh == 1 || a->kh == 3) && (a->kw == 1 || a->kw == 3)))
        return set_error("wgrad: taps per dim must be 1 or 3");
    if (a->layout == 1 && a->up_taps * cpad8(a->up_co) != cpad8(a->Co)) return set_error("wgrad: transposed-conv layout mismatch");
    return launch_wgrad_tc(a, (cudaStream_t)stream);
}

int e3b_wgrad_reduce_batched(const e3b_wgrad_args* args, int n, void* stream)
{
    if (!args || n <= 0) return set_error("wgrad_reduce_batched: no jobs");
    return launch_wgrad_reduce_batched(args, n, (cudaStream_t)stream);
}

}  // extern "C"
