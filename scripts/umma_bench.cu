// Micro-benchmark: cycles per tcgen05.mma (kind::tf32, M=128, cta_group::1) as a function of N and of the
// shared-memory operand layout.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../elektronn3_b200/csrc/common.cuh"
using namespace e3b;

struct Cfg { int N; int a_layout; uint32_t a_lbo, a_sbo; int b_layout; uint32_t b_lbo, b_sbo; int a_shift; int nacc; };

__device__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, int layout) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) bench(Cfg c, int iters, long long* out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_tf32(c.N, 0, 0);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
        uint64_t ad[8], bd[2];
        for (int j = 0; j < 8; j++) ad[j] = mk_desc(a0 + (uint32_t)(j * c.a_shift), c.a_lbo, c.a_sbo, c.a_layout);
        bd[0] = mk_desc(b0, c.b_lbo, c.b_sbo, c.b_layout); bd[1] = mk_desc(b0 + 16384, c.b_lbo, c.b_sbo, c.b_layout);
        uint32_t accs[8];
        for (int j = 0; j < 8; j++) accs[j] = tm + (uint32_t)((j % c.nacc) * c.N);
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) umma_tf32(accs[j], ad[j], bd[j & 1], idesc, 1u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main()
{
    long long* d; cudaMalloc(&d, 148 * sizeof(long long));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    auto run = [&](const char* name, Cfg c) {
        const int iters = 4000;
        bench<<<148, 128, 200 * 1024>>>(c, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-60s ERROR %s\n", name, cudaGetErrorString(e)); exit(1); }
        long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
        printf("N=%3d nacc=%d %-50s %.1f cycles/MMA (ideal %d)\n", c.N, c.nacc, name, avg / iters, c.N / 2);
    };
    // x / y tap shifts of the halo tile: A start += 16 B (one voxel in x) or 160 B (one row)
    for (int N : {32, 96, 192}) {
        const int nacc = 512 / N > 4 ? 4 : 512 / N;
        run("halo tile A, A start += 0", Cfg{N, 0, 2880, 160, 0, (uint32_t)N * 16, 128, 0, nacc});
        run("halo tile A, A start += 16 B per MMA (x taps)", Cfg{N, 0, 2880, 160, 0, (uint32_t)N * 16, 128, 16, nacc});
        run("halo tile A, A start += 32 B per MMA", Cfg{N, 0, 2880, 160, 0, (uint32_t)N * 16, 128, 32, nacc});
        run("halo tile A, A start += 160 B per MMA (y taps)", Cfg{N, 0, 2880, 160, 0, (uint32_t)N * 16, 128, 160, nacc});
        run("halo tile A (SBO=128 dense rows), A start += 16 B", Cfg{N, 0, 2880, 128, 0, (uint32_t)N * 16, 128, 16, nacc});
        run("halo tile A (SBO=256), A start += 16 B", Cfg{N, 0, 5760, 256, 0, (uint32_t)N * 16, 128, 16, nacc});
        run("halo tile A (SBO=256), A start += 0 B", Cfg{N, 0, 5760, 256, 0, (uint32_t)N * 16, 128, 0, nacc});
    }
    const int Ns[] = {32, 64, 128, 256};
    for (int N : Ns) {
        for (int nacc = 1; nacc <= 8 && nacc * N <= 512; nacc *= 2) {
            run("halo tile A (SBO=160, LBO=28800), A start += 2880 per MMA", Cfg{N, 0, 28800, 160, 0, (uint32_t)N * 16, 128, 2880, nacc});
        }
        run("halo tile A, same A start", Cfg{N, 0, 28800, 160, 0, (uint32_t)N * 16, 128, 0, 512 / N > 8 ? 8 : 512 / N});
        run("dense A (SBO=128, LBO=2048)", Cfg{N, 0, 2048, 128, 0, (uint32_t)N * 16, 128, 0, 512 / N > 8 ? 8 : 512 / N});
        run("sw128 A and B", Cfg{N, 2, 16, 1024, 2, 16, 1024, 0, 512 / N > 8 ? 8 : 512 / N});
    }
    return 0;
}
