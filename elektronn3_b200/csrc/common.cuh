// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05 (UMMA / TMEM) PTX wrappers.
// Hand-written inline PTX; no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define E3B_DEVINL __device__ __forceinline__

// ---------------------------------------------------------------- internal activation layout
// "QP" (quad-planar): fp32 tensor (N, Cq, D, H, W, 4) with Cq = ceil8(C) / 4.  Channel c lives in plane
// c / 4, lane c % 4.  Padding channels (c >= C) are kept at exactly 0.  One voxel of one plane is a
// 16-byte quad, so that (a) 8 consecutive W-voxels of a plane form one canonical no-swizzle UMMA core
// matrix (8 rows x 16 B) both as K-major operand (rows = voxels, K = channels: forward / dgrad) and as
// MN-major operand (K = voxels: wgrad), and (b) a spatial shift by one voxel is a +16 B shift of the
// shared-memory start address, which is what lets all 27 taps of a 3x3x3 stencil read ONE halo tile.
__host__ __device__ inline int e3b_cpad(int c) { return (c + 7) & ~7; }

// "QH" (operand layout): the tensors the forward / dgrad MMAs read -- network input, activations, pooled
// activations, conv-output gradients -- are stored as fp16 (N, Ch, D, H, W, 8) with Ch = ceil16(C) / 8:
// channel c lives in plane c / 8, lane c % 8; again one voxel of one plane is a 16-byte unit, so every
// statement above about core matrices and +16 B tap shifts holds unchanged, and one MMA (kind::f16, K = 16)
// consumes two planes.  fp16 carries the same 10 explicit mantissa bits as TF32 (the reference's GPU
// arithmetic): inside the fp16 normal range the stored operand is bit-for-bit the TF32-rounded value.
// Gradients are kept in range by a per-tensor power-of-two scale (e3b_norm_bwd_*), undone in the dgrad epilogue.
__host__ __device__ inline int e3b_cpad16(int c) { return (c + 15) & ~15; }

namespace e3b {

E3B_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
E3B_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
E3B_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
E3B_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

E3B_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
E3B_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
E3B_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost arrival (a bug) must surface as a launch error, never as a hung GPU box.
// try_wait suspends the thread for a HW-defined interval, so the bound is wall-clock (globaltimer).
#ifndef E3B_WAIT_TIMEOUT_NS
#define E3B_WAIT_TIMEOUT_NS 4000000000ull
#endif
E3B_DEVINL uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
E3B_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0xFF) == 0 && globaltimer_ns() - t0 > E3B_WAIT_TIMEOUT_NS) { __trap(); }
    }
}

// ---------------------------------------------------------------- TMA
E3B_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)m) : "memory");
}
E3B_DEVINL void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3,
                            int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
E3B_DEVINL void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// 1D bulk copy global -> shared (bytes % 16 == 0, both addresses 16 B aligned)
E3B_DEVINL void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"((uint64_t)gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
E3B_DEVINL void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "r"(ncols)
                 : "memory");
}
E3B_DEVINL void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
E3B_DEVINL void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
E3B_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
E3B_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.commit: arrive on an mbarrier when all previously issued MMAs of this thread have completed
E3B_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}

// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layouts, in 16-byte units:
//   K-major :  ((8,m),2) : ((1,SBO),LBO)   8 rows of a core matrix are 16 B apart, 8-row groups SBO
//              apart, the two 16-byte K halves of one tf32 MMA (K=8) LBO apart.
//   MN-major:  ((1,m),(8,k)) : ((X,SBO),(1,LBO))   4 consecutive M/N elements per 16 B, successive
//              16-byte M/N chunks SBO apart, the 8 K-steps of one MMA 16 B apart.
// bits: [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=0
E3B_DEVINL uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// Instruction descriptor for kind::tf32, fp32 accumulate, M = 128.
//   [4,6) c_format=1 (F32) | [7,10) a_format=2 (TF32) | [10,13) b_format=2 | [15] a_major | [16] b_major
//   [17,23) N>>3 | [24,29) M>>4
__host__ __device__ inline uint32_t umma_idesc_tf32(int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// Instruction descriptor for kind::f16 with fp16 operands (a_format = b_format = 0), fp32 accumulate, M = 128.
__host__ __device__ inline uint32_t umma_idesc_f16(int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}

E3B_DEVINL void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

E3B_DEVINL void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive 32-bit columns
E3B_DEVINL void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// TMEM -> registers without waiting: several loads are put in flight, then ONE tcgen05.wait::ld (tmem_ld_wait32)
E3B_DEVINL void tmem_ld16_raw(uint32_t taddr, uint32_t* r)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// wait for the loads, then pin the 32 destination registers behind the wait (the compiler must not consume them earlier)
E3B_DEVINL void tmem_ld_wait32(uint32_t* r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                      "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
    asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                      "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}

// 32-byte global store (two adjacent 16-byte units; the address must be 32-byte aligned): one full sector per lane
E3B_DEVINL void st_global_256(void* p, const uint4& a, const uint4& b)
{
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
                 "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

// round-to-nearest fp32 -> tf32 (10 explicit mantissa bits).  The tensor core ignores the low 13 bits of
// its fp32 operands (truncation, biased); tensors that feed an MMA are stored pre-rounded instead.
E3B_DEVINL float tf32_rn(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// 4 floats -> 4 halves (round to nearest even) packed in 8 bytes
E3B_DEVINL uint2 pack_half4(float a, float b, float c, float d) {
    const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
    uint2 r;
    r.x = *reinterpret_cast<const uint32_t*>(&lo);
    r.y = *reinterpret_cast<const uint32_t*>(&hi);
    return r;
}
E3B_DEVINL float4 unpack_half4(const uint2& u) {
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

E3B_DEVINL bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

E3B_DEVINL void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace e3b
