// Micro-test: may a tiled TMA load of a 2-byte element type start at an inner coordinate that is NOT a
// multiple of 8 elements (16 bytes)?  Decides whether the weight-gradient kernel can read its kw x-shifted
// operand boxes from ONE planar fp16 gradient copy (three loads at x0-1, x0, x0+1) instead of three
// materialised copies.  One case per process (an illegal instruction poisons the context):
//   tma_shift_test <x0> <swizzle 0|1>        prints OK / MISMATCH / the CUDA error
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_shift_test tma_shift_test.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../elektronn3_b200/csrc/common.cuh"
using namespace e3b;

static constexpr int W = 256, H = 32, BX = 64, BY = 8;

__global__ void k(const __grid_constant__ CUtensorMap m, int x0, int y0, __half* out)
{
    __shared__ __align__(1024) __half tile[BX * BY];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, BX * BY * 2);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(tile)), "l"((uint64_t)&m), "r"(smem_u32(&bar)), "r"(x0), "r"(y0) : "memory");
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BX * BY; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv)
{
    const int x0 = argc > 1 ? atoi(argv[1]) : 0;
    const int sw = argc > 2 ? atoi(argv[2]) : 0;
    std::vector<__half> h(W * H);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) h[y * W + x] = __float2half((float)((y * 37 + x) % 2048));
    __half *d, *o;
    cudaMalloc(&d, W * H * 2); cudaMalloc(&o, BX * BY * 2);
    cudaMemcpy(d, h.data(), W * H * 2, cudaMemcpyHostToDevice);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap m;
    cuuint64_t dims[2] = {W, H}, strides[1] = {W * 2};
    cuuint32_t box[2] = {BX, BY}, es[2] = {1, 1};
    CUresult r = ((Enc)fp)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("x0=%d sw=%d encode failed %d\n", x0, sw, (int)r); return 1; }
    const int y0 = 3;
    k<<<1, 128>>>(m, x0, y0, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("x0=%d sw=%d CUDA error: %s\n", x0, sw, cudaGetErrorString(e)); return 2; }
    std::vector<__half> res(BX * BY);
    cudaMemcpy(res.data(), o, BX * BY * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < BY; yy++)
        for (int xx = 0; xx < BX; xx++) {
            const int gx = x0 + xx, gy = y0 + yy;
            const float want = (gx >= 0 && gx < W && gy < H) ? (float)((gy * 37 + gx) % 2048) : 0.f;
            // 128-byte swizzle: 16-byte chunk index (xx/8) is XORed with the row index mod 8
            const int chunk = sw ? ((xx >> 3) ^ (yy & 7)) : (xx >> 3);
            const float got = __half2float(res[yy * BX + chunk * 8 + (xx & 7)]);
            if (got != want) { if (bad < 4) printf("  (%d,%d) got %g want %g\n", xx, yy, got, want); bad++; }
        }
    printf("x0=%d sw=%d %s (%d mismatches)\n", x0, sw, bad ? "MISMATCH" : "OK", bad);
    return 0;
}
