// "z-stacked" implicit-GEMM 3x3x3 convolution (forward and dgrad) for narrow outputs (pad16(Cout) <= 80) on
// tcgen05 tensor cores, sm_100a.  Same operator and operands as conv_tc.cu (elektronn3 `conv3`,
// models/unet.py:131-149, incl. the virtual torch.cat of unet.py:399 and the dgrad of both), different GEMM
// shape.
//
// Why: tcgen05.mma with both operands in shared memory fetches (128 + N) x 32 B per M=128, K=16 instruction;
// measured on B200 (scripts/umma_bench.cu) that costs 44 clocks at N = 32, 48 at N = 64, 56 at N = 96, 64 at
// N = 128, while the math takes N/2 clocks.  With N = Cout = 32 (the layers that hold most of the FLOPs of a
// UNet with start_filts=32) the halo-tile kernel of conv_tc.cu is operand-fetch bound at 16/44 = 36 % of the
// tensor peak.  Here the three z taps are STACKED IN N:
//
//     acc[v, (j, co)] += sum_{dy,dx,ci} x[z', v + (dy,dx), ci] * w[dz = 2 - j, dy, dx][ci][co]
//
// one MMA of N = 3*Cout per (dy, dx) tap and 16-channel chunk applies INPUT plane z' to the three OUTPUT
// planes z'+pd-2 .. z'+pd at once (N = 96: 48 math clocks of 56; N = 192: math bound).  The accumulators of
// successive output planes are consecutive column blocks of a TMEM ring, so the three contributions to an
// output plane (from input planes z-1, z, z+1) meet in the same TMEM columns and the tensor core does the
// sum: no partial results ever leave TMEM.  A CTA marches along z through a contiguous run of the
// (n, y tile, x tile, z) space: every input plane tile (8+2 x 16+2 voxels, all channels) is staged by TMA
// exactly once per run (no z halo re-reads), the complete weight image stays resident in shared memory.
//
// Roles: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc), warps 2..9 epilogue (two warpgroups taking
// alternate output planes).  The epilogue drains an output plane as soon as its last input plane has been
// applied (tcgen05.commit -> mbarrier), clears the block with tcgen05.st so that every MMA can accumulate,
// and hands it back.
//
// The single issuing thread is the scarce resource (measured, profiles/r02_umma_issue_queue_commit_wait.txt: one
// tcgen05.mma costs it ~46 cycles, one tcgen05.commit ~190 cycles of pipe time, one mbarrier wait on a complete
// phase ~116 cycles, and the MMA queue only hides ~4 MMAs), so its per-plane bookkeeping is kept minimal:
//  * a shared-memory stage holds ZB = 2 consecutive input planes (one TMA box, one full barrier, one release commit);
//  * the issuer never waits for TMEM blocks: the PRODUCER waits for the blocks a stage's MMAs will newly touch
//    (blk_free) and only then gives the stage's full barrier its second arrival, so "stage full" implies "blocks free";
//  * for narrow planes (NTW <= 48) a CTA runs TWO independent chains side by side (CHAINS = 2: two producers, two
//    issuers, one epilogue warpgroup each, half of the stages and half of the TMEM ring each, ONE shared weight image):
//    while one issuer is busy with waits / commits / bookkeeping the other one keeps the tensor pipe fed.  Each chain is
//    the same deterministic sequence of MMAs a single-chain CTA would issue for that part of the run.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace e3b {

static constexpr int kZsThreads1 = 320;           // one chain : TMA warp, MMA warp, two epilogue warpgroups
static constexpr int kZsThreads2 = 384;           // two chains: 2 TMA warps, 2 MMA warps, one epilogue warpgroup per chain
static constexpr int kZsTX = 8, kZsTY = 16;
static constexpr int kZsMaxBlocks = 32;          // TMEM ring blocks (512 / NTW, NTW >= 16)
static constexpr int kZsMaxStages = 8;
static constexpr int kZsMaxN = 80;               // widest output plane block (columns)

// kind::f16 MMA, always accumulating, descriptors given as (low, high) words: the issue loop advances the low
// words with 32-bit uniform adds
E3B_DEVINL void umma_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
}

// 16-value butterfly transpose-reduce over the warp: afterwards lane l < 16 holds in sm[0] / sq[0] the sums over all
// 32 lanes of column bitrev4(l)
E3B_DEVINL void zs_butterfly16(float* sm, float* sq, int lane)
{
#pragma unroll
    for (int step = 0; step < 4; step++) {
        const int keepn = 8 >> step;
        const int bit = 1 << step;
        const bool upper = (lane & bit) != 0;
#pragma unroll
        for (int j = 0; j < keepn; j++) {
            const float send_s = upper ? sm[j] : sm[j + keepn];
            const float send_q = upper ? sq[j] : sq[j + keepn];
            const float keep_s = upper ? sm[j + keepn] : sm[j];
            const float keep_q = upper ? sq[j + keepn] : sq[j];
            sm[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, bit);
            sq[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, bit);
        }
    }
    sm[0] += __shfl_xor_sync(0xffffffffu, sm[0], 16);
    sq[0] += __shfl_xor_sync(0xffffffffu, sq[0], 16);
}

struct ConvZsParams {
    int N, D, Do, Ho, Wo;        // input planes, output extents
    int pd, ph, pw;
    int HX, HY;                  // halo tile extents of one plane
    int tiles_x, tiles_y;
    long long total_L;           // (n, y tile, x tile) chains x Do output planes (pair: (n, y tile, x tile pair) x Do)
    int pair;                    // 1: the two chains of a CTA take x-ADJACENT tile columns over the same run of planes, so the
                                 //    128-byte lines both halo tiles touch cross HBM once (they meet in L2 within microseconds)
    int chunks0, chunks1;        // 16-channel K chunks of source 0 / 1
    int NTW, R, SA;              // columns per output plane (of one N tile), ring blocks, plane-tile stages
    int ZB;                      // input planes per stage
    int ntiles;                  // N tiles of NTW columns: the grid is cut into ntiles groups of CTAs, one tile each
    uint32_t a0_bytes, a_stage_bytes, w_bytes, w_piece_bytes;
    const float* bias; int n_bias;
    float* dst0; int cq0; float* dst1; int cq0_alloc, cq1_alloc;
    int relu, half_out;
    const float* out_scale; const float* w_unscale;
    double* stats; int Cstat;
    const uint8_t* wpk;
    uint32_t* dbg;               // host-mapped debug words or null
    int prof;                    // accumulate role timings into g_zs_prof
    int skip;                    // tuning aid (E3B_ZS_SKIP): 1 no global stores, 2 no statistics, 4 no MMAs, 8 streaming (evict-first) fp32 stores
};


// this warp's statistics of sample n -> global [N][Cstat][2] (fp64 atomics), accumulators cleared
E3B_DEVINL void zs_flush_stats(const ConvZsParams& p, int n, int co0, int lane, int bcol, bool reg_stats, float* rs, float* rq,
                               double* acc_s, double* acc_q)
{
    if (reg_stats) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            zs_butterfly16(rs + i * 16, rq + i * 16, lane);
            acc_s[i] += (double)rs[i * 16]; acc_q[i] += (double)rq[i * 16];
#pragma unroll
            for (int j = 0; j < 16; j++) { rs[i * 16 + j] = 0.f; rq[i * 16 + j] = 0.f; }
        }
    }
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const int ch = co0 + i * 16 + bcol;
        if (lane < 16 && i * 16 < p.NTW && ch < p.Cstat) {
            atomicAdd(p.stats + ((size_t)n * p.Cstat + ch) * 2, acc_s[i]);
            atomicAdd(p.stats + ((size_t)n * p.Cstat + ch) * 2 + 1, acc_q[i]);
        }
        acc_s[i] = 0.0; acc_q[i] = 0.0;
    }
}

struct ZsPiece { int n, y0, x0, za, zb; };      // output planes [za, zb) of one chain

E3B_DEVINL ZsPiece zs_piece(const ConvZsParams& p, long long L, long long L1, int chain_id)
{
    ZsPiece g;
    int chain = (int)(L / p.Do);
    g.za = (int)(L - (long long)chain * p.Do);
    if (p.pair) chain = 2 * chain + chain_id;
    const long long left = L1 - L;
    g.zb = (g.za + left < p.Do) ? (int)(g.za + left) : p.Do;
    const int xt = chain % p.tiles_x; chain /= p.tiles_x;
    const int yt = chain % p.tiles_y;
    g.n = chain / p.tiles_y;
    g.x0 = xt * kZsTX; g.y0 = yt * kZsTY;
    return g;
}
// input planes a piece needs (planes outside [0, D) are zero padding and are skipped altogether)
E3B_DEVINL int zs_in_lo(const ConvZsParams& p, const ZsPiece& g) { const int z = g.za - p.pd; return z < 0 ? 0 : z; }
E3B_DEVINL int zs_in_hi(const ConvZsParams& p, const ZsPiece& g) { const int z = g.zb + 1 - p.pd; return z > p.D - 1 ? p.D - 1 : z; }

E3B_DEVINL void tmem_st16_zero(uint32_t taddr)
{
    const uint32_t z = 0;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
        : "memory");
}
// bring-up aid (E3B_ZS_DEBUG): a wait that times out leaves (tag, progress) in host-mapped memory before the trap
E3B_DEVINL void zs_wait(uint64_t* bar, uint32_t parity, volatile uint32_t* dbg, uint32_t slot, uint32_t tag)
{
    if (!dbg) { mbar_wait(bar, parity); return; }
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0xFF) == 0 && globaltimer_ns() - t0 > 1000000000ull) {
            dbg[slot] = tag; __threadfence_system();
            const uint64_t t1 = globaltimer_ns();
            while (globaltimer_ns() - t1 < 500000000ull) { }
            __trap();
        }
    }
}

E3B_DEVINL void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Optional cycle accounting of the pipeline roles (E3B_ZS_PROF, scripts/zs_bench.py): summed over CTAs.
__device__ unsigned long long g_zs_prof[16];
#define ZP_T0(var) long long var = 0; if (p.prof) var = clock64()
#define ZP_ACC(slot, var) if (p.prof) { const long long now_ = clock64(); prof[slot] += (unsigned long long)(now_ - var); var = now_; }

template <int CHAINS>
__global__ void __launch_bounds__(CHAINS == 2 ? kZsThreads2 : kZsThreads1, 1)
conv_zs_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1, const ConvZsParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // roles: warps [0, CHAINS) TMA producers, [CHAINS, 2 CHAINS) MMA issuers, then 8 epilogue warps
    const int role = warp < CHAINS ? 0 : (warp < 2 * CHAINS ? 1 : 2);
    const int ew = warp - 2 * CHAINS;                                   // epilogue warp 0..7
    const int chain = role == 0 ? warp : (role == 1 ? warp - CHAINS : (CHAINS == 2 ? ew >> 2 : 0));
    // carve: [weights][plane-tile stages (SA per chain)][barriers]; p.SA and p.R are PER CHAIN
    uint8_t* w_base = smem;
    uint8_t* a_base = smem + p.w_bytes + (size_t)chain * p.SA * p.a_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.w_bytes + (size_t)CHAINS * p.SA * p.a_stage_bytes);
    uint64_t* a_full = bars + chain * (kZsMaxStages / CHAINS);                          // [kZsMaxStages]
    uint64_t* a_empty = bars + kZsMaxStages + chain * (kZsMaxStages / CHAINS);          // [kZsMaxStages]
    uint64_t* blk_full = bars + 2 * kZsMaxStages + chain * (kZsMaxBlocks / CHAINS);     // [kZsMaxBlocks]
    uint64_t* blk_free = bars + 2 * kZsMaxStages + kZsMaxBlocks + chain * (kZsMaxBlocks / CHAINS);   // [kZsMaxBlocks]
    uint64_t* w_full = bars + 2 * kZsMaxStages + 2 * kZsMaxBlocks;                      // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
    __shared__ __align__(16) float bias_s[kZsMaxN];

    const int nchunks = p.chunks0 + p.chunks1;
    const int N3 = 3 * p.NTW;

    if (threadIdx.x == 0) {
        // a_full: the TMA transaction (arrive.expect_tx) + the producer's "output blocks are free" arrival
        for (int c = 0; c < CHAINS; c++) {
            for (int i = 0; i < p.SA; i++) {
                mbar_init(&bars[c * (kZsMaxStages / CHAINS) + i], 2);
                mbar_init(&bars[kZsMaxStages + c * (kZsMaxStages / CHAINS) + i], 1);
            }
            for (int i = 0; i < p.R; i++) {
                mbar_init(&bars[2 * kZsMaxStages + c * (kZsMaxBlocks / CHAINS) + i], 1);
                mbar_init(&bars[2 * kZsMaxStages + kZsMaxBlocks + c * (kZsMaxBlocks / CHAINS) + i], 4);
            }
        }
        mbar_init(w_full, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmap0);
        if (p.chunks1) tma_prefetch_desc(&tmap1);
    }
    if (warp == CHAINS) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot + (uint32_t)chain * (512u / CHAINS);      // this chain's half of the columns

    // N tiles (wide outputs): CTA group `ntile` computes output columns [ntile * NTW, (ntile + 1) * NTW) of the whole volume
    const int G = (int)gridDim.x / p.ntiles;
    const int ntile = (int)blockIdx.x / G, bloc = (int)blockIdx.x % G;
    // this CTA's contiguous run of the linear (chain, output plane) space, cut into one contiguous part per chain
    const long long C0 = p.total_L * bloc / G, C1 = p.total_L * (bloc + 1) / G;
    const long long L0 = p.pair ? C0 : C0 + (C1 - C0) * chain / CHAINS, L1 = p.pair ? C1 : C0 + (C1 - C0) * (chain + 1) / CHAINS;

    if (role == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            if (chain == 0) {                        // the weight image is shared by the chains
                mbar_arrive_expect_tx(w_full, p.w_bytes);
                const uint8_t* wsrc = p.wpk + (size_t)ntile * p.w_bytes;
                for (uint32_t o = 0; o < p.w_bytes; o += p.w_piece_bytes) bulk_load_1d(w_base + o, wsrc + o, p.w_piece_bytes, w_full);
            }
            uint32_t sa = 0, pa = 0;
            uint32_t ws = 0, free_par = 0;           // ring slot of the next output plane to acquire, its barriers' parities
            unsigned long long prof[16] = {0};
            ZP_T0(t_all); ZP_T0(t);
            for (long long L = L0; L < L1;) {
                const ZsPiece g = zs_piece(p, L, L1, chain);
                const int zlo = zs_in_lo(p, g), zhi = zs_in_hi(p, g);
                int next_wait = g.za;
                for (int zp = zlo; zp <= zhi; zp += p.ZB) {
                    const int zlast = zp + p.ZB - 1 < zhi ? zp + p.ZB - 1 : zhi;       // last input plane of this stage
                    ZP_ACC(15, t);
                    zs_wait(&a_empty[sa], pa ^ 1, p.dbg, 16 * blockIdx.x + 0, 0x100000u | (sa << 8) | (uint32_t)zp);
                    ZP_ACC(1, t);
                    // (the box always spans ZB planes: a plane beyond the piece is loaded but not used, one beyond the
                    // volume is zero-filled; either way the transaction is the full stage)
                    mbar_arrive_expect_tx(&a_full[sa], p.a_stage_bytes);
                    uint8_t* st = a_base + (size_t)sa * p.a_stage_bytes;
                    tma_load_5d(st, &tmap0, &a_full[sa], (g.x0 - p.pw) * 4, g.y0 - p.ph, zp, 0, g.n);
                    if (p.chunks1) tma_load_5d(st + p.a0_bytes, &tmap1, &a_full[sa], (g.x0 - p.pw) * 4, g.y0 - p.ph, zp, 0, g.n);
                    // TMEM blocks of the output planes this stage's MMAs touch for the first time: wait here, on the
                    // producer's time, so that the issuer never does
                    int ohi = zlast + p.pd; if (ohi > g.zb - 1) ohi = g.zb - 1;
                    for (; next_wait <= ohi; next_wait++) {
                        zs_wait(&blk_free[ws], (free_par >> ws) & 1u, p.dbg, 16 * blockIdx.x + 1, 0x200000u | (ws << 8) | (uint32_t)zp);
                        free_par ^= 1u << ws;
                        if (++ws == (uint32_t)p.R) ws = 0;
                    }
                    ZP_ACC(3, t);
                    mbar_arrive(&a_full[sa]);
                    if (++sa == (uint32_t)p.SA) { sa = 0; pa ^= 1; }
                }
                L += g.zb - g.za;
            }
            ZP_ACC(0, t_all);
            if (p.prof) { for (int i = 0; i < 2; i++) atomicAdd(&g_zs_prof[i], prof[i]); atomicAdd(&g_zs_prof[3], prof[3]); }
        }
    } else if (role == 1) {
        // ===================== MMA issuer =====================
        // The whole warp runs the loop with uniform control flow, so that all descriptor arithmetic stays on the
        // uniform datapath (32-bit adds on the low descriptor word: the address field cannot carry); one elected
        // lane issues the MMAs and the commits.
        const bool leader = elect_one();
        const uint32_t idesc1 = umma_idesc_f16(p.NTW, 0, 0), idesc2 = umma_idesc_f16(2 * p.NTW, 0, 0),
                       idesc3 = umma_idesc_f16(3 * p.NTW, 0, 0);
        const uint32_t plane16 = (uint32_t)(p.HX * p.HY);               // one 8-channel plane of one z plane, 16-byte units
        const uint32_t ZB = (uint32_t)p.ZB;
        const uint64_t a_tmpl = umma_desc(0, ZB * plane16 * 16u, (uint32_t)(p.HX * 16));
        const uint64_t b_tmpl = umma_desc(0, (uint32_t)(N3 * 16), 128);
        const uint32_t a_hi = (uint32_t)(a_tmpl >> 32), b_hi = (uint32_t)(b_tmpl >> 32);
        const uint32_t a_lo0 = (uint32_t)a_tmpl + (smem_u32(a_base) >> 4), b_lo0 = (uint32_t)b_tmpl + (smem_u32(w_base) >> 4);
        const uint32_t a_stage16 = p.a_stage_bytes >> 4;
        const uint32_t HX = (uint32_t)p.HX, NTW = (uint32_t)p.NTW, R = (uint32_t)p.R;
        const uint32_t chunk_step_a = 2u * ZB * plane16, tap_step_b = (uint32_t)(2 * N3);
        zs_wait(w_full, 0, p.dbg, 16 * blockIdx.x + 14, 0x500000u);
        tc_fence_after();
        uint32_t sa = 0, pa = 0;
        uint32_t ws = 0;                         // ring slot of the next output plane to enter the window (== slot of the next piece's first plane at piece end)
        unsigned long long prof[16] = {0};
        ZP_T0(t_all); ZP_T0(t);
        for (long long L = L0; L < L1;) {
            const ZsPiece g = zs_piece(p, L, L1, chain);
            const int zlo = zs_in_lo(p, g), zhi = zs_in_hi(p, g);
            int next_in = g.za, next_commit = g.za, olo_prev = g.za;
            uint32_t cs = ws, s_lo = ws;         // slots of next_commit / of the lowest output plane fed by the current input plane
            for (int zp0 = zlo; zp0 <= zhi; zp0 += p.ZB) {
                const int zlast = zp0 + p.ZB - 1 < zhi ? zp0 + p.ZB - 1 : zhi;
                ZP_ACC(15, t);
                // one wait per stage: the TMA has landed AND (second arrival, by the producer) the TMEM blocks of every
                // output plane this stage touches are free
                zs_wait(&a_full[sa], pa, p.dbg, 16 * blockIdx.x + 2, 0x300000u | (sa << 8) | (uint32_t)zp0);
                ZP_ACC(4, t);
                tc_fence_after();
                for (int zp = zp0; zp <= zlast; zp++) {
                    // output planes this input plane feeds: z = zp + pd - dz, dz = 0..2, inside the piece
                    int olo = zp + p.pd - 2; if (olo < g.za) olo = g.za;
                    int ohi = zp + p.pd; if (ohi > g.zb - 1) ohi = g.zb - 1;
                    if (olo != olo_prev) { olo_prev = olo; if (++s_lo == R) s_lo = 0; }      // olo advances by at most one per plane
                    for (; next_in <= ohi; next_in++) { if (++ws == R) ws = 0; }
                    const int nb = ohi - olo + 1;
                    const uint32_t j_lo = (uint32_t)(olo - (zp + p.pd - 2));
                    int n1 = (int)(R - s_lo); if (n1 > nb) n1 = nb;
                    const int n2 = nb - n1;
                    const uint32_t id1 = n1 == 3 ? idesc3 : (n1 == 2 ? idesc2 : idesc1);
                    const uint32_t id2 = n2 == 2 ? idesc2 : idesc1;
                    const uint32_t acc1 = tmem_base + s_lo * NTW;
                    uint32_t a_lo = a_lo0 + sa * a_stage16 + (uint32_t)(zp - zp0) * plane16;
                    uint32_t b_lo = b_lo0 + j_lo * NTW;
                    const uint32_t b2_off = (uint32_t)n1 * NTW;
                    ZP_ACC(12, t);
                    if (nb > 0 && !(p.skip & 4)) {
                        if (n2 == 0) {
                            for (int c = 0; c < nchunks; c++) {
                                if (leader) {
#pragma unroll
                                    for (int ty = 0; ty < 3; ty++)
#pragma unroll
                                        for (int tx = 0; tx < 3; tx++)
                                            umma_f16_lohi(acc1, a_lo + (uint32_t)ty * HX + (uint32_t)tx, a_hi,
                                                          b_lo + (uint32_t)(ty * 3 + tx) * tap_step_b, b_hi, id1);
                                }
                                a_lo += chunk_step_a; b_lo += 9u * tap_step_b;
                            }
                        } else {
                            // the three blocks wrap around the end of the TMEM ring: two MMAs per tap
                            for (int c = 0; c < nchunks; c++) {
                                if (leader) {
#pragma unroll
                                    for (int ty = 0; ty < 3; ty++)
#pragma unroll
                                        for (int tx = 0; tx < 3; tx++) {
                                            const uint32_t al = a_lo + (uint32_t)ty * HX + (uint32_t)tx;
                                            const uint32_t bl = b_lo + (uint32_t)(ty * 3 + tx) * tap_step_b;
                                            umma_f16_lohi(acc1, al, a_hi, bl, b_hi, id1);
                                            umma_f16_lohi(tmem_base, al, a_hi, bl + b2_off, b_hi, id2);
                                        }
                                }
                                a_lo += chunk_step_a; b_lo += 9u * tap_step_b;
                            }
                        }
                    }
                    // (one commit site per barrier kind only: a second, `else if (leader)` copy of a commit was once
                    // compiled to an unpredicated UTCBAR that the non-elected lanes executed too -> double arrival)
                    ZP_ACC(13, t);
                    // output planes whose last contributing input plane (min(z - pd + 2, D - 1)) has now been issued
                    for (; next_commit < g.zb; next_commit++) {
                        int last = next_commit - p.pd + 2; if (last > p.D - 1) last = p.D - 1;
                        if (last > zp) break;
                        if (leader) umma_commit(&blk_full[cs]);
                        if (++cs == R) cs = 0;
                    }
                    ZP_ACC(14, t);
                }
                if (leader) umma_commit(&a_empty[sa]);          // the stage (all its planes) goes back to the producer
                __syncwarp();
                ZP_ACC(5, t);
                if (++sa == (uint32_t)p.SA) { sa = 0; pa ^= 1; }
            }
            L += g.zb - g.za;
        }
        ZP_ACC(2, t_all);
        if (p.prof && leader) { atomicAdd(&g_zs_prof[2], prof[2]); atomicAdd(&g_zs_prof[4], prof[4]); atomicAdd(&g_zs_prof[5], prof[5]); for (int i = 12; i < 16; i++) atomicAdd(&g_zs_prof[i], prof[i]); }
    } else {
        // ===================== epilogue: two warpgroups =====================
        // one chain : they take alternate output planes (EG = 2);  two chains: one warpgroup per chain, every plane (EG = 1)
        constexpr uint32_t EG = CHAINS == 2 ? 1u : 2u;
        const int eg = CHAINS == 2 ? 0 : (ew >> 2);      // drains the output planes whose running count is eg (mod EG)
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // GEMM row inside the tile
        const int ry = row >> 3, rx = row & 7;
        const int etid = threadIdx.x - 64 * CHAINS;      // 0..255
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        // butterfly transpose-reduce leaves column (bit-reversed low nibble of the lane) in lanes 0..15
        const int bcol = ((lane & 1) << 3) | ((lane & 2) << 1) | ((lane & 4) >> 1) | ((lane & 8) >> 3);
        const int co0 = ntile * p.NTW;          // first output column (channel) of this CTA's N tile
        for (int i = etid; i < kZsMaxN; i += 256) bias_s[i] = (p.bias && co0 + i < p.n_bias) ? p.bias[co0 + i] : 0.f;
        // clear the whole ring once, then hand every block to the issuer
        if (eg == 0) {
            for (int c = 0; c < p.R * p.NTW; c += 16) tmem_st16_zero(lane_base + (uint32_t)c);
            tmem_wait_st();
            tc_fence_before();
        }
        named_bar_sync(1, 256);                 // bias_s visible to the epilogue warps
        if (eg == 0 && lane == 0) for (int b = 0; b < p.R; b++) mbar_arrive(&blk_free[b]);

        const float oscale = (p.out_scale ? __ldg(p.out_scale) : 1.f) * (p.w_unscale ? __ldg(p.w_unscale) : 1.f);
        const bool reg_stats = p.stats != nullptr && p.NTW <= 32;     // per-thread partial sums over a whole run
        float rs[32], rq[32];
#pragma unroll
        for (int i = 0; i < 32; i++) { rs[i] = 0.f; rq[i] = 0.f; }
        double acc_s[5], acc_q[5];              // this lane's channel (cg + bcol) of every 16-column group
#pragma unroll
        for (int i = 0; i < 5; i++) { acc_s[i] = 0.0; acc_q[i] = 0.0; }
        int cur_n = -1;
        uint32_t s = 0, full_par = 0, cnt = 0;  // (EG = 2: R is even, a ring slot always meets the same warpgroup)
        const uint32_t R = (uint32_t)p.R;
        const size_t cstride = (size_t)p.Do * p.Ho * p.Wo;            // 16-byte units between channel planes of the output
        unsigned long long prof[16] = {0};
        ZP_T0(t_all); ZP_T0(t);
        for (long long L = L0; L < L1;) {
            const ZsPiece g = zs_piece(p, L, L1, chain);
            if (p.stats && g.n != cur_n) {
                if (cur_n >= 0) zs_flush_stats(p, cur_n, co0, lane, bcol, reg_stats, rs, rq, acc_s, acc_q);
                cur_n = g.n;
            }
            const int y = g.y0 + ry, x = g.x0 + rx;
            const bool valid = (y < p.Ho) && (x < p.Wo);
            const bool store = valid && !(p.skip & 1);
            // 16-byte-unit offset of (n, plane 0, z = 0, y, x) in dst0 / dst1
            const size_t vox = (size_t)y * p.Wo + x;
            const size_t o0 = (size_t)g.n * p.cq0_alloc * cstride + vox, o1 = (size_t)g.n * p.cq1_alloc * cstride + vox;
            for (int z = g.za; z < g.zb; z++, cnt++) {
                if (EG == 2 && (cnt & 1u) != (uint32_t)eg) { if (++s == R) s = 0; continue; }
                ZP_ACC(15, t);
                zs_wait(&blk_full[s], (full_par >> s) & 1u, p.dbg, 16 * blockIdx.x + 3 + q, 0x400000u | (s << 8) | (uint32_t)z);
                ZP_ACC(7, t);
                full_par ^= 1u << s;
                tc_fence_after();
                const uint32_t blk = lane_base + s * (uint32_t)p.NTW;
                const size_t zoff = (size_t)z * p.Ho * p.Wo;
#pragma unroll
                for (int ip = 0; ip < 3; ip++) {
                  // 32 columns at a time: both TMEM loads in flight before the one wait
                  const int cg0 = ip * 32;
                  if (cg0 < p.NTW) {
                    const bool two = cg0 + 16 < p.NTW;
                    uint32_t r[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) r[j] = 0u;
                    tmem_ld16_raw(blk + (uint32_t)cg0, r);
                    if (two) tmem_ld16_raw(blk + (uint32_t)cg0 + 16u, r + 16);
                    tmem_ld_wait32(r);
                    tmem_st16_zero(blk + (uint32_t)cg0);             // the block is clear again for its next output plane
                    if (two) tmem_st16_zero(blk + (uint32_t)cg0 + 16u);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                      const int i = ip * 2 + h;
                      const int cg = i * 16;
                      if (h == 0 || two) {
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; j++) v[j] = __uint_as_float(r[h * 16 + j]);
#pragma unroll
                        for (int j4 = 0; j4 < 4; j4++) {
                            const float4 b = *reinterpret_cast<const float4*>(&bias_s[cg + j4 * 4]);
                            v[j4 * 4] = fmaf(v[j4 * 4], oscale, b.x); v[j4 * 4 + 1] = fmaf(v[j4 * 4 + 1], oscale, b.y);
                            v[j4 * 4 + 2] = fmaf(v[j4 * 4 + 2], oscale, b.z); v[j4 * 4 + 3] = fmaf(v[j4 * 4 + 3], oscale, b.w);
                        }
                        if (p.relu) {
#pragma unroll
                            for (int j = 0; j < 16; j++) v[j] = fmaxf(v[j], 0.f);
                        }
                        ZP_ACC(8, t);
                        if (store) {
                            if (p.half_out) {
                                // operand tensor (QH): 16 columns = two 16-byte units (planes cg/8, cg/8 + 1)
#pragma unroll
                                for (int j8 = 0; j8 < 2; j8++) {
                                    const int hpl = ((co0 + cg) >> 3) + j8;
                                    if (hpl < p.cq0_alloc) {
                                        const uint2 lo = pack_half4(v[j8 * 8], v[j8 * 8 + 1], v[j8 * 8 + 2], v[j8 * 8 + 3]);
                                        const uint2 hi = pack_half4(v[j8 * 8 + 4], v[j8 * 8 + 5], v[j8 * 8 + 6], v[j8 * 8 + 7]);
                                        reinterpret_cast<uint4*>(p.dst0)[o0 + (size_t)hpl * cstride + zoff] = make_uint4(lo.x, lo.y, hi.x, hi.y);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int j4 = 0; j4 < 4; j4++) {
                                    const int cq = ((co0 + cg) >> 2) + j4;
                                    const float4 val = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                                    if (cq < p.cq0) {
                                        if (cq < p.cq0_alloc) {
                                            float4* q = reinterpret_cast<float4*>(p.dst0) + (o0 + (size_t)cq * cstride + zoff);
                                            if (p.skip & 8) __stcs(q, val); else *q = val;
                                        }
                                    } else if (p.dst1 != nullptr && cq - p.cq0 < p.cq1_alloc) {
                                        reinterpret_cast<float4*>(p.dst1)[o1 + (size_t)(cq - p.cq0) * cstride + zoff] = val;
                                    }
                                }
                            }
                        }
                        ZP_ACC(9, t);
                        if (p.stats && !(p.skip & 2)) {
                            if (reg_stats) {
                                if (i < 2 && valid) {
#pragma unroll
                                    for (int j = 0; j < 16; j++) { rs[(i & 1) * 16 + j] += v[j]; rq[(i & 1) * 16 + j] = fmaf(v[j], v[j], rq[(i & 1) * 16 + j]); }
                                }
                            } else {
                                float sm[16], sq[16];
#pragma unroll
                                for (int j = 0; j < 16; j++) { sm[j] = valid ? v[j] : 0.f; sq[j] = sm[j] * sm[j]; }
                                zs_butterfly16(sm, sq, lane);
                                acc_s[i] += (double)sm[0];
                                acc_q[i] += (double)sq[0];
                            }
                        }
                        ZP_ACC(10, t);
                      }
                    }
                  }
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&blk_free[s]);
                if (++s == R) s = 0;
                ZP_ACC(11, t);
            }
            L += g.zb - g.za;
        }
        if (p.stats && cur_n >= 0) zs_flush_stats(p, cur_n, co0, lane, bcol, reg_stats, rs, rq, acc_s, acc_q);
        ZP_ACC(6, t_all);
        if (p.prof && (ew & 3) == 0 && lane == 0) for (int i = 6; i < 12; i++) atomicAdd(&g_zs_prof[i], prof[i]);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == CHAINS) { tc_fence_after(); tmem_dealloc(*tmem_slot, 512); }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int zs_max_dyn_smem()
{
    static int v = 0;
    if (!v) {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, conv_zs_kernel<1>) != cudaSuccess) return 0;
        v = 227 * 1024 - (int)fa.sharedSizeBytes;
    }
    return v;
}

static const size_t kZsBarrierBytes = (size_t)(2 * kZsMaxStages + 2 * kZsMaxBlocks + 1) * 8 + 16;

// Shared-memory / TMEM plan: resident weight image + per chain SA stages of ZB input-plane tiles and R ring blocks.
struct ZsPlan { int chains, zb, sa, r; uint32_t w_bytes, a_stage_bytes; };

static int env_int(const char* name)
{
    const char* e = getenv(name);
    return e ? atoi(e) : 0;
}

// ok = false: the weights (plus a minimal ring of stages) do not fit.
static bool zs_plan(int C0, int C1, int n_total, ZsPlan* out)
{
    const int chunks = cpad16(C0) / 16 + (C1 > 0 ? cpad16(C1) / 16 : 0);
    const size_t wb = (size_t)chunks * 9 * 2 * (3 * n_total) * 16;
    const size_t ab = (size_t)(kZsTX + 2) * (kZsTY + 2) * 16 * 2 * chunks;      // one input plane tile, all channels
    const size_t budget = 227 * 1024 - 2048 - kZsBarrierBytes;       // static smem (bias) and alignment slack: 1 KB each
    if (wb + 2 * ab > budget) return false;
    const size_t room = budget - wb;
    static int force_chains = -1, force_zb = -1;                      // tuning / tests: E3B_ZS_CHAINS, E3B_ZS_ZB
    if (force_chains < 0) { force_chains = env_int("E3B_ZS_CHAINS"); force_zb = env_int("E3B_ZS_ZB"); }
    ZsPlan pl;
    pl.w_bytes = (uint32_t)wb;
    // two chains: narrow planes only (each chain gets half of the 512 TMEM columns), at least two stages per chain
    for (int chains = 2; chains >= 1; chains--) {
        if (force_chains && chains != force_chains) continue;
        if (chains == 2 && n_total > 48) continue;
        int r = (512 / chains) / n_total;
        if (r > kZsMaxBlocks / chains) r = kZsMaxBlocks / chains;
        if (chains == 1) r &= ~1;                                     // the two warpgroups own alternate ring slots
        for (int zb = 2; zb >= 1; zb--) {
            if (force_zb && zb != force_zb) continue;
            if (r < zb + 4) continue;                                 // zb + 2 blocks accumulate, the rest is slack for the epilogue
            const size_t stage = ab * zb;
            size_t sa = room / (stage * chains);
            const size_t cap = zb == 2 ? 4 : (size_t)kZsMaxStages / chains;
            if (sa > cap) sa = cap;
            const size_t need = (chains == 1 && zb == 1) ? 3 : 2;      // single-plane stages of a lone chain: at least three
            if (sa < need) continue;
            pl.chains = chains; pl.zb = zb; pl.sa = (int)sa; pl.r = r; pl.a_stage_bytes = (uint32_t)stage;
            if (out) *out = pl;
            return true;
        }
    }
    return false;
}

// columns per N tile: the whole padded output width when its weight image fits shared memory, else tiles of 32 columns
// (N = 3 * 32 = 96 per MMA); 0: this convolution is not served by the z-stacked kernel
int conv_zs_ntile(int C0, int C1, int n_total)
{
    if (n_total % 16 || n_total < 16 || C0 <= 0 || C1 < 0) return 0;
    if (n_total <= kZsMaxN && zs_plan(C0, C1, n_total, nullptr)) return n_total;
    if (n_total % 32 == 0 && n_total <= 512 && zs_plan(C0, C1, 32, nullptr)) return 32;
    return 0;
}

int conv_zs_supported(int C0, int C1, int n_total, int kd, int kh, int kw, int scatter)
{
    static int disabled = -1;
    if (disabled < 0) { const char* e = getenv("E3B_CONV_VARIANT"); disabled = (e && atoi(e) == 0) ? 1 : 0; }
    if (disabled) return 0;
    if (scatter || kd != 3 || kh != 3 || kw != 3) return 0;
    return conv_zs_ntile(C0, C1, n_total) > 0 ? 1 : 0;
}

int conv_zs_prof_read(unsigned long long* out16, int reset)
{
    cudaError_t e = cudaMemcpyFromSymbol(out16, g_zs_prof, sizeof(unsigned long long) * 16);
    if (e != cudaSuccess) return set_error("conv_zs_prof_read: %s", cudaGetErrorString(e));
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_zs_prof, z, sizeof(z));
    }
    return 0;
}

static uint32_t* g_zs_dbg_host = nullptr;
int conv_zs_debug_read(uint32_t* out, int n)
{
    if (!g_zs_dbg_host) return set_error("zs debug buffer not allocated (E3B_ZS_DEBUG unset?)");
    for (int i = 0; i < n && i < 16 * 148; i++) out[i] = g_zs_dbg_host[i];
    return 0;
}

int launch_conv_zs(const e3b_conv_args* a, cudaStream_t stream)
{
    ConvZsParams p;
    memset(&p, 0, sizeof(p));
    if (a->kd != 3 || a->kh != 3 || a->kw != 3 || a->scatter) return set_error("conv(z-stacked): 3x3x3 convolutions only");
    p.N = a->N; p.D = a->D;
    p.Do = a->D + 2 * a->pd - 2; p.Ho = a->H + 2 * a->ph - 2; p.Wo = a->W + 2 * a->pw - 2;
    if (p.Do <= 0 || p.Ho <= 0 || p.Wo <= 0) return set_error("conv: empty output");
    p.pd = a->pd; p.ph = a->ph; p.pw = a->pw;
    p.NTW = conv_zs_ntile(a->C0, a->src1 ? a->C1 : 0, a->n_total);
    if (p.NTW <= 0) return set_error("conv(z-stacked): output width %d / input channels %d not supported", a->n_total, a->C0);
    p.ntiles = a->n_total / p.NTW;
    p.HX = kZsTX + 2; p.HY = kZsTY + 2;
    p.tiles_x = (p.Wo + kZsTX - 1) / kZsTX; p.tiles_y = (p.Ho + kZsTY - 1) / kZsTY;
    p.total_L = (long long)a->N * p.tiles_x * p.tiles_y * p.Do;
    p.chunks0 = cpad16(a->C0) / 16;
    p.chunks1 = a->src1 ? cpad16(a->C1) / 16 : 0;
    ZsPlan pl;
    if (!zs_plan(a->C0, a->src1 ? a->C1 : 0, p.NTW, &pl)) return set_error("conv(z-stacked): weights do not fit shared memory");
    p.SA = pl.sa; p.ZB = pl.zb; p.R = pl.r; p.w_bytes = pl.w_bytes; p.a_stage_bytes = pl.a_stage_bytes;
    if (const char* e = getenv("E3B_ZS_SA")) { const int v = atoi(e); if (v >= 2 && v < p.SA) p.SA = v; }           // tuning / tests
    p.a0_bytes = (uint32_t)(p.HX * p.HY * 16 * 2 * p.chunks0 * p.ZB);
    p.w_piece_bytes = (uint32_t)(2 * 3 * p.NTW * 16);                 // one (chunk, tap) image
    p.bias = a->bias; p.n_bias = a->n_bias;
    p.dst0 = reinterpret_cast<float*>(a->dst0); p.dst1 = a->dst1;
    p.cq0_alloc = a->half_out ? cpad16(a->Cd0) / 8 : cpad8(a->Cd0) / 4;
    p.cq1_alloc = a->dst1 ? cpad8(a->Cd1) / 4 : 0;
    p.cq0 = a->dst1 ? p.cq0_alloc : (1 << 30);
    p.relu = a->relu; p.half_out = a->half_out; p.out_scale = a->out_scale; p.w_unscale = a->w_unscale;
    if (a->half_out && a->dst1) return set_error("conv: the fp16 operand output has a single destination");
    if (a->half_out && a->stats) return set_error("conv: statistics are taken from an fp32 output");
    p.stats = a->stats; p.Cstat = a->stats_channels;
    p.wpk = reinterpret_cast<const uint8_t*>(a->wpk);

    CUtensorMap m0, m1;
    int rc = make_qp_tensor_map(&m0, reinterpret_cast<const float*>(a->src0), a->N, p.chunks0 * 2, a->D, a->H, a->W, a->D, a->H,
                                a->W, p.HX, p.HY, p.ZB, p.chunks0 * 2);
    if (rc) return rc;
    if (a->src1) {
        const float* v1 = reinterpret_cast<const float*>(a->src1) + (((size_t)a->off1_d * a->H1 + a->off1_h) * a->W1 + a->off1_w) * 4;
        rc = make_qp_tensor_map(&m1, v1, a->N, p.chunks1 * 2, a->D, a->H, a->W, a->D1, a->H1, a->W1, p.HX, p.HY, p.ZB, p.chunks1 * 2);
        if (rc) return rc;
    } else {
        m1 = m0;
    }
    const size_t smem = (size_t)p.w_bytes + (size_t)pl.chains * p.SA * p.a_stage_bytes + kZsBarrierBytes + 1024;
    static bool configured[kMaxDevices] = {false};
    if (!configured[current_device()]) {
        const int max_dyn = zs_max_dyn_smem();
        if (max_dyn <= 0) return set_error("conv(z-stacked): cudaFuncGetAttributes failed");
        cudaError_t e = cudaFuncSetAttribute(conv_zs_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_zs_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
        if (e != cudaSuccess) return set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured[current_device()] = true;
    }
    if (p.stats) {
        cudaError_t e = cudaMemsetAsync(p.stats, 0, sizeof(double) * 2 * (size_t)a->N * p.Cstat, stream);
        if (e != cudaSuccess) return set_error("conv: stats memset: %s", cudaGetErrorString(e));
    }
    if (const char* e = getenv("E3B_ZS_SKIP")) p.skip = atoi(e);
    p.prof = getenv("E3B_ZS_PROF") != nullptr;
    if (getenv("E3B_ZS_DEBUG")) {
        static uint32_t* dbg_dev = nullptr;
        if (!g_zs_dbg_host) {
            if (cudaHostAlloc(&g_zs_dbg_host, 16 * 148 * sizeof(uint32_t), cudaHostAllocMapped) != cudaSuccess) return set_error("zs debug alloc");
            cudaHostGetDevicePointer(&dbg_dev, g_zs_dbg_host, 0);
        }
        memset(g_zs_dbg_host, 0, 16 * 148 * sizeof(uint32_t));
        p.dbg = dbg_dev;
    }
    p.pair = (pl.chains == 2 && p.tiles_x % 2 == 0 && !getenv("E3B_ZS_NO_PAIR")) ? 1 : 0;
    if (p.pair) p.total_L = (long long)a->N * (p.tiles_x / 2) * p.tiles_y * p.Do;
    long long grid = num_sms() / p.ntiles;                            // CTAs per N tile
    if (grid > p.total_L) grid = p.total_L;
    if (grid < 1) grid = 1;
    if (const char* e = getenv("E3B_ZS_GRID")) { const int gcap = atoi(e); if (gcap > 0 && gcap < grid) grid = gcap; }   // tests: long runs on small volumes
    grid *= p.ntiles;
    if (pl.chains == 2) conv_zs_kernel<2><<<(int)grid, kZsThreads2, smem, stream>>>(m0, m1, p);
    else conv_zs_kernel<1><<<(int)grid, kZsThreads1, smem, stream>>>(m0, m1, p);
    return check_launch("conv_zs");
}

}  // namespace e3b
