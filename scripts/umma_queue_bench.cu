// Micro-benchmark of what surrounds tcgen05.mma in the conv kernels' issue loops (B200):
//  (1) issue-time profile: clock after each of 64 back-to-back MMAs (N = 96) -> how deep the MMA queue is before the
//      issuing thread blocks, and the latency of the first MMA on an idle pipe;
//  (2) cost of an mbarrier try_wait on an ALREADY COMPLETE phase, executed by 32 lanes vs by one lane;
//  (3) cost of a tcgen05.commit between MMAs;
//  (4) MMA rate with a second warp streaming bulk copies into the same CTA's shared memory (TMA write traffic).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_queue_bench umma_queue_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../elektronn3_b200/csrc/common.cuh"
using namespace e3b;

__device__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// mode 0: issue profile; 1: wait cost; 2: MMA batches of `batch` separated by a commit; 3: MMAs + concurrent bulk copies
__global__ void __launch_bounds__(128, 1) bench(int mode, int N, int batch, int copy_bytes, const uint8_t* gsrc, long long* out, int mn)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar2, cbar[4];
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1); mbar_init(&bar2, 1);
        for (int i = 0; i < 4; i++) mbar_init(&cbar[i], 1);
        fence_barrier_init(); stop = 0;
    }
    if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // mn = 1: both operands MN-major (the weight-gradient kernel: K = voxels 16 B apart, 8-channel groups one tile plane apart)
    const uint32_t idesc = umma_idesc_f16(N, mn, mn);
    const uint64_t ad = mn ? mk_desc(smem_u32(smem), 128, 544) : mk_desc(smem_u32(smem), 2880, 160);
    const uint64_t bd = mn ? mk_desc(smem_u32(smem + 64 * 1024), 128, 512) : mk_desc(smem_u32(smem + 64 * 1024), (uint32_t)N * 16, 128);
    if (mode == 0 && threadIdx.x == 0) {
        long long t[65];
        t[0] = clock64();
#pragma unroll
        for (int i = 0; i < 64; i++) { umma_f16(tm + (i & 3) * N, ad, bd, idesc, 1u); t[i + 1] = clock64(); }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long tend = clock64();
        for (int i = 0; i < 64; i++) out[i] = t[i + 1] - t[0];
        out[64] = tend - t[0];
        // first-MMA latency on an idle pipe: one MMA + commit + wait
        const long long a = clock64();
        umma_f16(tm, ad, bd, idesc, 1u);
        umma_commit(&bar2);
        mbar_wait(&bar2, 0);
        out[65] = clock64() - a;
    }
    if (mode == 1 && warp == 0) {
        if (lane == 0) mbar_arrive(&bar);           // phase 0 complete
        __syncwarp();
        long long a = clock64();
        for (int i = 0; i < 64; i++) { mbar_wait(&bar, 0); __syncwarp(); }
        long long b = clock64();
        for (int i = 0; i < 64; i++) { if (lane == 0) mbar_wait(&bar, 0); __syncwarp(); }
        long long c = clock64();
        if (lane == 0) { out[0] = (b - a) / 64; out[1] = (c - b) / 64; }
    }
    if (mode == 2 && threadIdx.x == 0) {
        const int iters = 2048;
        uint32_t ph = 0;
        const long long a = clock64();
        for (int i = 0; i < iters; i += batch) {
            for (int j = 0; j < batch; j++) umma_f16(tm + (j & 3) * N, ad, bd, idesc, 1u);
            umma_commit(&bar2);                      // (arrivals pile up on a count-1 barrier: phases just advance)
            (void)ph;
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        out[0] = (clock64() - a) / iters;
    }
    if (mode == 3) {
        if (threadIdx.x == 0) {
            const int iters = 4096;
            const long long a = clock64();
            for (int i = 0; i < iters; i++) umma_f16(tm + (i & 3) * N, ad, bd, idesc, 1u);
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            out[0] = (clock64() - a) / iters;
            stop = 1;
        } else if (warp == 1 && lane == 0 && copy_bytes > 0) {
            // stream bulk copies into a 4-slot ring above the operands (no consumer: only the write traffic matters)
            uint32_t s = 0, par[4] = {0, 0, 0, 0};
            bool pending[4] = {false, false, false, false};
            long long n = 0;
            const long long a = clock64();
            while (!stop) {
                if (pending[s]) { mbar_wait(&cbar[s], par[s]); par[s] ^= 1; }
                mbar_arrive_expect_tx(&cbar[s], copy_bytes);
                bulk_load_1d(smem + 96 * 1024 + s * 16384, gsrc + ((n * 16384) & (64 * 1024 * 1024 - 1)), copy_bytes, &cbar[s]);
                pending[s] = true;
                s = (s + 1) & 3; n++;
            }
            out[1] = n * copy_bytes * 1000 / (clock64() - a);    // bytes per 1000 cycles
            for (int k = 0; k < 4; k++) if (pending[k]) mbar_wait(&cbar[k], par[k]);   // nothing in flight at exit
        }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main()
{
    long long* d; cudaMalloc(&d, 128 * sizeof(long long));
    uint8_t* g; cudaMalloc(&g, 64 * 1024 * 1024 + 65536); cudaMemset(g, 0, 64 * 1024 * 1024 + 65536);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long h[128];
    auto run = [&](int mode, int N, int batch, int cb, int grid, int mn = 0) {
        cudaMemset(d, 0, sizeof(h));
        bench<<<grid, 128, 200 * 1024>>>(mode, N, batch, cb, g, d, mn);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d ERROR %s\n", mode, cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    };
    for (int N : {32, 96}) {
        run(0, N, 0, 0, 1); run(0, N, 0, 0, 1);
        printf("N=%d issue profile (cycles since start after MMA i):", N);
        for (int i = 0; i < 64; i++) printf(" %lld", h[i]);
        printf("\n  all 64 complete after %lld cycles; one MMA + commit + wait on an idle pipe: %lld cycles\n", h[64], h[65]);
    }
    run(1, 96, 0, 0, 1);
    printf("mbarrier wait on a complete phase: %lld cycles with 32 lanes polling, %lld with one lane (+__syncwarp)\n", h[0], h[1]);
    for (int batch : {1, 2, 4, 9, 18, 36}) {
        run(2, 96, batch, 0, 1);
        printf("N=96, tcgen05.commit after every %2d MMAs: %lld cycles/MMA\n", batch, h[0]);
    }
    for (int cb : {0, 2048, 8192, 16384}) {
        run(3, 96, 0, cb, 148);
        printf("N=96 with concurrent bulk copies of %5d B (148 CTAs): %lld cycles/MMA, copy rate %.1f B/cycle/SM\n", cb, h[0], h[1] / 1000.0);
        run(3, 64, 0, cb, 148);
        printf("N=64 with concurrent bulk copies of %5d B (148 CTAs): %lld cycles/MMA, copy rate %.1f B/cycle/SM\n", cb, h[0], h[1] / 1000.0);
    }
    for (int N : {48, 96, 128, 192, 256}) {
        run(3, N, 0, 0, 148, 1);
        const long long mnc = h[0];
        run(3, N, 0, 0, 148, 0);
        printf("N=%3d: %lld cycles/MMA with MN-major operands (LBO 128, SBO 544/512), %lld K-major\n", N, mnc, h[0]);
    }
    return 0;
}
