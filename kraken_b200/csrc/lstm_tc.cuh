// lstm_tc.cuh - LSTM recurrence on tcgen05 tensor cores (sm_100a): k_lstm_rec_tc<GL> for hidden sizes 129..256 (clusters of 8 CTAs)
// and k_lstm_rec_tc_small<NLS> for hidden sizes <= 32 (one CTA).  DESIGN.md 4.2 has the measured history (v1..v9) and what bounds it.
//
// Same contract as k_lstm_rec (kernels.cuh): per-pixel gate pre-activations gx in, hidden states out (fp32 and / or the fp16 operand
// planes of a tensor-core consumer), packed-sequence semantics, one direction per blockIdx.y.  The CUDA-core kernel is at the
// 3-register-FFMA issue limit, so W_hh . h_{t-1} moves to the tensor cores:
//
//   cluster of 8 CTAs = 2 groups x GL sequences of one direction for all time steps; CTA r owns unit slots [32r, 32r+32) (padded, any
//   hid <= 256), i.e. 128 gate rows ordered unit-major (row = 4*slot + gate: a unit's gates sit in 4 adjacent TMEM lanes).
//
//   Operand precision: fp16 pairs with a power-of-two scale on the second term give 22 significand bits,
//       x = x1 + x2s * 2^-11,   x1 = fp16(x),   x2s = fp16((x - x1) * 2^11)        (|error| <= 2^-23 |x|)
//   (a first attempt with bf16 x bf16 / fp16 x bf16 mixed-format MMAs faulted with "illegal instruction": A and B of one kind::f16
//   MMA must share a format.)
//   A = W1 and W2s planes of the CTA's 128 gate rows x K = 256, resident in TENSOR MEMORY for the whole kernel (row = TMEM lane, two
//       fp16 per 32-bit column, 2 x 128 columns; written once with tcgen05.st).  With A in shared memory every MMA re-read its 4 KB
//       slice: 128 KB per step, measured 1158 cycles to issue the 32 MMAs.
//   B (per group, double buffered): h_{t-1} planes as swizzle-free core matrices [k-chunk of 8 unit slots][row = (warp half, plane,
//       line)][8 slots] fp16; every 16-byte row is written REMOTELY by the CTA that computed it (st.async from registers, byte-counted
//       by the destination's mbarrier: no cluster barrier, no "buffer free" handshake thanks to the double buffer).
//   D: per group and step 32 MMAs   W1 x [h1|h2s]   and   W2s x [h1|..]   (M128 x N 2GL x K16) into FOUR TMEM accumulator chains
//       (product x K half), interleaved by the group's one issuer warp.  Every fp32 accumulator element sees only 8 accumulations.
//       pre = D1[:, h1] + 2^-11 (D1[:, h2s] + D2[:, h1])                               (dropped: W2*h2 ~ 2^-24)
//   epilogue: 16 warps = 2 groups x 4 TMEM lane quarters x 2: tcgen05.ld -> + gx (prefetched a step ahead) -> SFU sigmoid / tanh
//       (ex2.approx + rcp.approx, ~1e-7) -> gates regrouped by quad shuffles -> fp32 cell update -> 8-lane gather of the line's
//       16-byte operand rows -> st.async to the 8 CTAs; h_t to HBM (row stores).  The two groups alternate on the tensor pipe.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gemm_tc.cuh"

namespace kb {
namespace ltc {

using namespace kb::tc;

constexpr int NG_MAX = 4;                                // independent line groups per cluster (own issuer, barriers, accumulators, epilogue warps): 2, 3 or 4
constexpr int LCS = 8;                                   // cluster size
constexpr int A_PLANE_ELEMS = 128 * 256;                 // fp16 elements of one W plane of a CTA: [128 gate rows][K = 256], row-major
constexpr int EW = 4;                                    // epilogue warps per TMEM lane quarter: 2 per group (latency hiding: v3 with 1 was 2x slower)
// warps 0..NG-1: MMA issuers (warp 0 also owns the TMEM allocation); then 4 EH epilogue warps per group (EH = 2: two warps per TMEM lane
// quarter, each half of the group's lines; EH = 1: one warp per quarter with all of them - what lets four groups fit 1024 threads)
constexpr int lthreads(int ng, int eh = 2) { return (ng + 4 * eh * ng) * 32; }
constexpr int LTHREADS = lthreads(2);
constexpr int TM_COLS = 512;
template <int GL, int NG = 2, int EH = 2> struct ClusterCfg {   // GL = lines per group (8 or 16), NG groups: 16 / 24 / 32 lines per cluster
    static constexpr int NL = NG * GL;                   // lines per cluster
    static constexpr int LPW = GL / EH;                  // lines per epilogue warp (EH warps per TMEM lane quarter and group)
    static constexpr int NT = LPW / 4;                   // 4x4 gate transposes (= cells) per thread and step
    static constexpr int CH_B = 2 * GL * 16;             // bytes of one k-chunk (8 unit slots) of a group: rows [warp half][plane][line] x 16 B
    static constexpr int B_BUF_B = 32 * CH_B;            // bytes per (group, buffer): K = 256 = 32 chunks
    static constexpr int SRC_B = 4 * CH_B;               // bytes one source CTA contributes to a buffer (its 4 k-chunks)
    static constexpr int HALF_B = 4 * SRC_B;             // bytes the four source CTAs of a K half deliver
    static constexpr int SMEM_BYTES = NG * 2 * B_BUF_B + 512 + 1024;
    static constexpr int N1 = 2 * GL;                    // N of both products (W2s x the same operand rows; only its h1 columns are read)
    static constexpr int GSTRIDE = 4 * N1;               // TMEM columns per group: D1a @0, D1b @N1, D2a @2 N1, D2b @3 N1
    static constexpr int TM_A0 = NG * GSTRIDE;           // A: W1 @TM_A0 (128 columns), W2s @TM_A0 + 128
    static constexpr int THREADS = lthreads(NG, EH);
    static_assert(LPW == 4 || LPW == 8, "an epilogue warp reads 4 or 8 accumulator columns per product");
    static_assert(NG * GSTRIDE + 256 <= 512, "accumulators + W_hh planes must fit the 512 TMEM columns");
};

struct LstmTcParams {
    const float *gx; const uint16_t *wpk; float *out; const int *lens;
    __half *out_hi, *out_lo;         // optional fp16 operand planes of the output for a tensor-core consumer (out may then be NULL)
    int nseq, T, hid, dirs, U;
    int q2; long long s_outer, s_inner, step;
    int alt;                         // 1: the groups take turns on the tensor pipe (see the issuer loop)
    int lpc;                         // lines per cluster (<= NL): group 0 takes ceil(lpc / 2) of them, group 1 the rest; fewer lines = fewer bytes through DSMEM per step
    int dbg; long long *dbgbuf;      // KB_LSTM_DBG: bit 0 = clock64 stamps of steps 100..103 of cluster 0 into dbgbuf[step][group][12]
};

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ uint32_t mapa32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// 16 bytes from registers into the shared memory of a CTA of the cluster, byte-counted by an mbarrier in that CTA
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint4 v, uint32_t rmbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rmbar) : "memory");
}
// K-major operand WITHOUT swizzle: core matrices (8 rows x 16 bytes, 128 contiguous bytes) `lbo` bytes apart along K and `sbo`
// bytes apart along M/N.  The h operand uses it because one 16-byte store = 8 unit slots of one line = one row of a core matrix.
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100); layout type 0 = no swizzle
    return d;
}
// non-suspending poll (mbarrier.test_wait): try_wait may park the warp for an implementation-defined time
__device__ __forceinline__ void mbar_wait_poll(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "PW_%=:\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra PD_%=;\n\t"
        "bra PW_%=;\n\t"
        "PD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// gate non-linearities on the SFU: ex2.approx + rcp, absolute error ~1e-7 (the CUDA-core kernel keeps expf/tanhf)
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }   // MUFU.RCP, no IEEE fix-up path
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_fast(1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.f, sigmoid_fast(2.f * x), -1.f); }
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// 4x4 transpose across the four lanes of a quad: in: v[i] = this lane's value for item i; out: v[j] = lane (quad base + j)'s value
// for item (lane & 3).  Two butterfly rounds (swap within 2x2 blocks, then swap the blocks), four shuffles, everything a SELECT on
// a lane-bit predicate with static register indices.  (A first version selected v[lane ^ k] with nested ternaries: ptxas turned
// them into 4-way divergent branches with BSSY/BSYNC - 1300 cycles per step in the KB_LSTM_DBG timeline.)
__device__ __forceinline__ void quad_transpose4(float (&v)[4], int gq) {
    const bool b0 = gq & 1, b1 = gq & 2;
    {
        const float sa = b0 ? v[0] : v[1], sb = b0 ? v[2] : v[3];
        const float ra = __shfl_xor_sync(0xffffffffu, sa, 1), rb = __shfl_xor_sync(0xffffffffu, sb, 1);
        v[0] = b0 ? ra : v[0]; v[1] = b0 ? v[1] : ra; v[2] = b0 ? rb : v[2]; v[3] = b0 ? v[3] : rb;
    }
    {
        const float sa = b1 ? v[0] : v[2], sb = b1 ? v[1] : v[3];
        const float ra = __shfl_xor_sync(0xffffffffu, sa, 2), rb = __shfl_xor_sync(0xffffffffu, sb, 2);
        v[0] = b1 ? ra : v[0]; v[2] = b1 ? v[2] : ra; v[1] = b1 ? rb : v[1]; v[3] = b1 ? v[3] : rb;
    }
}
// The eight lanes that share (lane & 3) hold one line's h for unit slots jq = lane >> 2 = 0..7 as (h1, h2s) halves.  After three
// butterfly rounds (7 shuffles) EVERY one of them holds the line's two complete 16-byte operand rows [8 unit slots] of the h1 and
// the h2s plane - so lane jq can hand them to CTA jq of the cluster.
__device__ __forceinline__ void gather_rows8(__half h1, __half h2, int lane, uint4 &row1, uint4 &row2) {
    const uint32_t my = (uint32_t)__half_as_ushort(h1) | ((uint32_t)__half_as_ushort(h2) << 16);
    const uint32_t ot = __shfl_xor_sync(0xffffffffu, my, 4);
    const bool b0 = lane & 4, b1 = lane & 8, b2 = lane & 16;
    const uint32_t lo = b0 ? ot : my, hi = b0 ? my : ot;                     // unit slots (jq & ~1), (jq | 1)
    const uint32_t p1 = (lo & 0xffffu) | (hi << 16), p2 = (lo >> 16) | (hi & 0xffff0000u);
    const uint32_t o1 = __shfl_xor_sync(0xffffffffu, p1, 8), o2 = __shfl_xor_sync(0xffffffffu, p2, 8);
    const uint32_t a0 = b1 ? o1 : p1, a1 = b1 ? p1 : o1, c0 = b1 ? o2 : p2, c1 = b1 ? p2 : o2;
    const uint32_t ra0 = __shfl_xor_sync(0xffffffffu, a0, 16), ra1 = __shfl_xor_sync(0xffffffffu, a1, 16);
    const uint32_t rc0 = __shfl_xor_sync(0xffffffffu, c0, 16), rc1 = __shfl_xor_sync(0xffffffffu, c1, 16);
    row1 = b2 ? make_uint4(ra0, ra1, a0, a1) : make_uint4(a0, a1, ra0, ra1);
    row2 = b2 ? make_uint4(rc0, rc1, c0, c1) : make_uint4(c0, c1, rc0, rc1);
}

// Synchronisation of one (group, time step), all through mbarriers, no CTA- or cluster-wide barrier in the loop:
//   b_half[g][buffer][K half]  tx-count barrier: the 4 KB of h_{s-1} planes the four source CTAs of a K half store (st.async from
//                              registers, 16 bytes per store) into this CTA's operand buffer.  Waited on by the group's issuer, which
//                              re-arms it for the refill two steps later once it has passed it (nobody can send that refill before
//                              receiving this CTA's h_s, which needs this step's MMAs).
//   acc_free[g]                4 EH arrivals: the group's epilogue warps have read the accumulators of the previous step.
//   mma_done[g]                tcgen05.commit of the group's issuer: accumulators complete.  Also read (not consumed) by the OTHER
//                              group's issuer: the groups alternate on the tensor pipe.
template <int GL, int NG, int EH = 2>
__global__ void __launch_bounds__(lthreads(NG, EH), 1) k_lstm_rec_tc(LstmTcParams p) {
    using Cfg = ClusterCfg<GL, NG, EH>;
    constexpr int N_ISSUE = NG, LTHREADS = Cfg::THREADS;
    constexpr int NL = Cfg::NL, CH_B = Cfg::CH_B, B_BUF_B = Cfg::B_BUF_B, LPW = Cfg::LPW, NT = Cfg::NT;
    constexpr int TM_A0 = Cfg::TM_A0, GSTRIDE = Cfg::GSTRIDE, N1 = Cfg::N1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_align1024(smem_raw);
    uint8_t *sB = smem;                                   // [group][buffer][k-chunk][row = (warp half, plane, line)][8 unit slots] fp16, no swizzle
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + NG * 2 * B_BUF_B);
    uint64_t *b_half = bars /* [group][buffer][K half] */, *mma_done = bars + 4 * NG /* [group] */, *acc_free = bars + 5 * NG /* [group] */;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 6 * NG);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int chunk = blockIdx.x / LCS, dir = blockIdx.y;
    const int hid = p.hid, GC = p.dirs * 4 * hid, OC = p.dirs * hid;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4 * NG; ++i) mbar_init(&b_half[i], 1);
        for (int g = 0; g < NG; ++g) { mbar_init(&mma_done[g], 1); mbar_init(&acc_free[g], EH * 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int g = 0; g < NG; ++g)
            for (int hf = 0; hf < 2; ++hf)       // buffer 1 of each group receives h_0 at the end of step 0
                mbar_expect_tx(&b_half[(g * 2 + 1) * 2 + hf], (uint32_t)(4 * 4 * 32 * max(0, min((p.lpc + NG - 1) / NG, p.lpc - g * ((p.lpc + NG - 1) / NG)))));
    }
    for (int i = threadIdx.x; i < NG * 2 * B_BUF_B / 16; i += LTHREADS) reinterpret_cast<uint4 *>(sB)[i] = make_uint4(0, 0, 0, 0);   // h_{-1} = 0
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic zero fill -> visible to UMMA reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= N_ISSUE && warp < N_ISSUE + 4) {
        // W_hh planes -> tensor memory: this thread's gate row (TMEM lane 32q + lane), K pairs packed low|high per column
        const int m = 32 * (warp & 3) + lane;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(p.wpk) + (((size_t)dir * LCS + rank) * 2 * A_PLANE_ELEMS) / 2 + (size_t)m * 128;
#pragma unroll 1
        for (int pl = 0; pl < 2; ++pl)
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t r[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)pl * (A_PLANE_ELEMS / 2) + c0 + i));
                    r[i] = v.x; r[i + 1] = v.y; r[i + 2] = v.z; r[i + 3] = v.w;
                }
                tmem_st32(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(TM_A0 + pl * 128 + c0), r);
            }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");

    // lines of this cluster: [chunk * lpc, chunk * lpc + lpc), ceil(lpc / NG) (<= GL) per group, the last group(s) take what is left
    const int lpc = p.lpc, lper = (lpc + NG - 1) / NG;
    int maxlen = 0;                                       // uniform across the cluster (both groups run the same number of steps)
    for (int lb = 0; lb < lpc; ++lb) {
        const int q = chunk * lpc + lb;
        if (q < p.nseq) maxlen = max(maxlen, p.lens ? min(max(p.lens[q], 0), p.T) : p.T);
    }
    const bool dbg_cta = (p.dbg & 1) && blockIdx.x == 0 && blockIdx.y == 0;

    if (warp < N_ISSUE) {
        // ===================== one MMA issuer per group.  Per step 32 MMAs M128 x N(2 GL) x K16 in four accumulator chains (product W1 /
        // W2s x K half), interleaved so that dependent MMAs are two issues apart.  Measured (KB_LSTM_DBG): a tiny MMA costs ~17 cycles
        // of tensor pipe whatever its N (the 128 x N fp32 accumulator read-modify-write, not the math), so the 32 MMAs of a group and
        // step take ~550 cycles however many threads issue them - eight issuer warps (one per chain and group) bought nothing and their
        // mbarrier polling competed with the epilogue warps' shuffles for the MIO pipe.
        const int g = warp;
        const uint32_t id1 = idesc_f16(0, 0, 128, N1);
        const uint32_t half_bytes = (uint32_t)(4 * 4 * 32 * max(0, min(lper, lpc - g * lper)));     // 4 source CTAs x 4 k-chunks x (h1 + h2s row) per real line slot
        for (int s = 0; s < maxlen; ++s) {
            const int cur = s & 1;
            const long long d_w0 = dbg_cta ? clock64() : 0;
            if (s > 0) mbar_wait(&acc_free[g], (uint32_t)((s - 1) & 1));    // complete long before h_{s-1} can arrive
            // Alternation: with two groups the tensor pipe is handed from one group to the other - g0(s), g1(s), g0(s+1), ... (the other
            // group's mma_done barrier is only READ here: parity wait, no arrival).  Left alone two groups fall into lock-step (both
            // start at step 0, and whoever is behind catches up while the other queues on the shared pipe / DSMEM port) and idle
            // together: 0.326 vs 0.311 ms on cfg2.  With three groups the same hand-over serialises them into 3 x 1370 cycles per step
            // (0.467 ms); free-running they settle at 0.356 ms (p.alt = 0).
            if (p.alt) {
                if (g == 0) { if (s > 0) mbar_wait(&mma_done[NG - 1], (uint32_t)((s - 1) & 1)); }
                else mbar_wait(&mma_done[g - 1], (uint32_t)(s & 1));
            }
            long long d_w1 = 0;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                uint64_t *bh = &b_half[(g * 2 + cur) * 2 + half];
                if (s > 0) {                                                  // default acquire.cta: an acquire.cluster wait costs a CCTL.IVALL per step
                    if (p.dbg & 2) mbar_wait_poll(bh, (uint32_t)(((s - 1) >> 1) & 1));
                    else mbar_wait(bh, (uint32_t)(((s - 1) >> 1) & 1));
                }
                if (half == 1) d_w1 = dbg_cta ? clock64() : 0;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.async writes (generic proxy) -> UMMA reads (async proxy); measured: ~60 cycles per step
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t b0 = smem_u32(sB + (g * 2 + cur) * B_BUF_B) + (uint32_t)(half * 16 * CH_B);
                    const uint32_t dacc = tmem_base + (uint32_t)(g * GSTRIDE + half * N1);
                    const uint32_t abase = tmem_base + (uint32_t)(TM_A0 + half * 64);      // K16 step = 8 columns of fp16 pairs
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {                          // K16 step = k-chunks 2j, 2j+1
                        const uint64_t bd = umma_desc_nosw(b0 + (uint32_t)(jj * 2 * CH_B), (uint32_t)CH_B, 128u);
                        umma_f16_ts(dacc, abase + (uint32_t)(jj * 8), bd, id1, jj ? 1u : 0u);                               // W1  x [h1 | h2s]
                        umma_f16_ts(dacc + (uint32_t)(2 * N1), abase + 128u + (uint32_t)(jj * 8), bd, id1, jj ? 1u : 0u);   // W2s x [h1 | ..]
                    }
                    if (s + 2 < maxlen) mbar_expect_tx(bh, half_bytes);       // refilled during step s+1 (nobody can send that before receiving our h_s)
                }
                __syncwarp();
            }
            if (elect_one()) {
                umma_commit(&mma_done[g]);
                if (dbg_cta && g < 2 && s >= 100 && s < 104) {
                    long long *d = p.dbgbuf + ((s - 100) * 2 + g) * 12;
                    d[0] = d_w0; d[1] = d_w1; d[2] = clock64();
                }
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue warps 2..17: quarter q = warp & 3; sub-warp sw = (warp-2) >> 2: group g = sw >> 1, half sw2 = sw & 1.
        // A warp owns LPW lines of its group x the 8 unit slots of its TMEM lane quarter and never synchronises with another warp:
        // gate values are regrouped by quad shuffles, h leaves as st.async straight from registers.
        const int q = warp & 3;                            // TMEM lane quarter: unit slots 8q .. 8q+7
        const int sw = (warp - N_ISSUE) >> 2, g = sw / EH, sw2 = sw % EH;
        const int jq = lane >> 2, gate = lane & 3;         // unit slot within the quarter, gate of this thread's TMEM row
        const int slot = 8 * q + jq, u = (int)rank * p.U + slot;
        const bool uvalid = slot < p.U && u < hid;
        // after the transposes this thread updates unit `u` of lines (first line of the warp) + 4 t + gate
        const int nlg = max(0, min(lper, lpc - g * lper));    // real line slots of this group
        const int line0 = chunk * lpc + g * lper + LPW * sw2;
        int clen[NT]; long long ooff[NT]; bool cval[NT]; float cst[NT];
        const long long ostride = (long long)(dir ? -1 : 1) * p.step * OC;
#pragma unroll
        bool slot_live[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int ql = line0 + 4 * t + gate;
            slot_live[t] = LPW * sw2 + 4 * t + gate < nlg;
            cval[t] = slot_live[t] && ql < p.nseq && uvalid;
            clen[t] = (slot_live[t] && ql < p.nseq) ? min(max(p.lens ? p.lens[ql] : p.T, 0), p.T) : 0;
            const int qq = ql < p.nseq ? ql : 0;
            const long long cbase = (long long)(qq / p.q2) * p.s_outer + (long long)(qq % p.q2) * p.s_inner;
            cst[t] = 0.f;
            if (cval[t]) for (int tt = clen[t]; tt < p.T; ++tt) {
                const size_t o = (size_t)(cbase + (long long)tt * p.step) * OC + dir * hid + u;
                if (p.out) p.out[o] = 0.f;
                if (p.out_hi) { p.out_hi[o] = __float2half_rn(0.f); p.out_lo[o] = __float2half_rn(0.f); }
            }
            ooff[t] = (cbase + (long long)(dir ? max(clen[t] - 1, 0) : 0) * p.step) * OC + dir * hid + u;
        }
        // gx of this thread's TMEM row (gate of unit u) for its LPW lines: running pointers, fetched one time step ahead
        int glen[LPW]; const float *gptr[LPW]; float gxn[LPW];
        const long long gstride = (long long)(dir ? -1 : 1) * p.step * GC;
#pragma unroll
        for (int i = 0; i < LPW; ++i) {
            const int ql = line0 + i;
            const bool v = LPW * sw2 + i < nlg && ql < p.nseq && uvalid;
            glen[i] = v ? min(max(p.lens ? p.lens[ql] : p.T, 0), p.T) : 0;
            const int qq = ql < p.nseq ? ql : 0;
            const long long gb = (long long)(qq / p.q2) * p.s_outer + (long long)(qq % p.q2) * p.s_inner;
            const int t0 = dir ? max(glen[i] - 1, 0) : 0;
            gptr[i] = p.gx + (size_t)(gb + (long long)t0 * p.step) * GC + (size_t)dir * 4 * hid + (size_t)(uvalid ? u : 0) * 4 + gate;
            gxn[i] = 0 < glen[i] ? __ldg(gptr[i]) : 0.f;
            gptr[i] += gstride;
        }
        // hand-off: after the 8-lane gather every lane holds the complete operand rows of its line; lane jq stores them into CTA jq.
        // This CTA's rows of the operand of every destination: k-chunk (4 rank + q), row (warp half, plane, line)
        const uint32_t row_off = ((uint32_t)rank * 4u + (uint32_t)q) * (uint32_t)CH_B + (uint32_t)((sw2 * 2 * LPW + gate) * 16);
        const uint32_t dstB = mapa32(smem_u32(sB), (uint32_t)jq) + row_off;
        const uint32_t dstBar = mapa32(smem_u32(b_half), (uint32_t)jq) + (uint32_t)(rank >> 2) * 8u;     // our K half's barrier at CTA jq
        // this warp's accumulator columns: rows of the operand = [warp half][plane][line]
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * GSTRIDE + 2 * LPW * sw2);
        // gate non-linearity without divergence: sigmoid(x) for i, f, o; tanh(x) = 2*sigmoid(2x) - 1 for the candidate gate
        const float act_k = gate == 2 ? 2.f : 1.f;
        const bool dbg_w = dbg_cta && g < 2 && q == 0 && sw2 == 0 && lane == 0;
        // whole 8-slot rows of this quarter are real units and 16-byte aligned in the output planes (hid 256: always)
        const bool vec_planes = p.out_hi && 8 * q + 7 < p.U && (int)rank * p.U + 8 * q + 7 < hid && (OC & 7) == 0 &&
                                ((dir * hid + (int)rank * p.U + 8 * q) & 7) == 0 &&
                                (reinterpret_cast<uintptr_t>(p.out_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out_lo) & 15) == 0;

        for (int s = 0; s < maxlen; ++s) {
            const int nxt = (s + 1) & 1;
            float gxv[LPW];
#pragma unroll
            for (int i = 0; i < LPW; ++i) gxv[i] = gxn[i];
            const long long e_top = dbg_w ? clock64() : 0;
            if (p.dbg & 4) mbar_wait_poll(&mma_done[g], (uint32_t)(s & 1));
            else mbar_wait(&mma_done[g], (uint32_t)(s & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long e_done = dbg_w ? clock64() : 0;
            // group columns: D1a = W1 x rows (K half 0) @0, D1b (K half 1) @N1, D2a = W2s x rows @2 N1, D2b @3 N1; within the warp's
            // 2 LPW rows: [h1 lines | h2s lines]
            uint32_t m0[LPW], m1[LPW], c0[LPW], c1[LPW], c2[LPW], c3[LPW];
            if (LPW == 4) {
                tmem_ld4_nowait(lane_base + 0, m0);        tmem_ld4_nowait(lane_base + N1, m1);
                tmem_ld4_nowait(lane_base + LPW, c0);      tmem_ld4_nowait(lane_base + N1 + LPW, c1);
                tmem_ld4_nowait(lane_base + 2 * N1, c2);   tmem_ld4_nowait(lane_base + 3 * N1, c3);
            } else {
                tmem_ld8_nowait(lane_base + 0, m0);        tmem_ld8_nowait(lane_base + N1, m1);
                tmem_ld8_nowait(lane_base + LPW, c0);      tmem_ld8_nowait(lane_base + N1 + LPW, c1);
                tmem_ld8_nowait(lane_base + 2 * N1, c2);   tmem_ld8_nowait(lane_base + 3 * N1, c3);
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[g]);       // the issuers may overwrite the accumulators
            const long long e_ld = dbg_w ? clock64() : 0;
            float av[LPW];
#pragma unroll
            for (int i = 0; i < LPW; ++i) {
                const float main_ = __uint_as_float(m0[i]) + __uint_as_float(m1[i]);
                const float corr = (__uint_as_float(c0[i]) + __uint_as_float(c1[i])) + (__uint_as_float(c2[i]) + __uint_as_float(c3[i]));
                const float pre = (main_ + corr * (1.f / X2_SCALE)) + gxv[i];
                av[i] = fmaf(sigmoid_fast(pre * act_k), act_k, 1.f - act_k);
            }
            const long long e_act = dbg_w ? clock64() : 0;
            float hv[NT]; __half hh1[NT], hh2[NT]; bool wr[NT];
            uint4 row1[NT], row2[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                float gt[4] = {av[4 * t], av[4 * t + 1], av[4 * t + 2], av[4 * t + 3]};
                quad_transpose4(gt, gate);                 // now: i, f, g, o of unit u for line 4 t + gate
                const bool live = s < clen[t];
                float h = 0.f;                             // finished / padding cells feed zeros (never used again)
                if (cval[t] && live) {
                    cst[t] = gt[1] * cst[t] + gt[0] * gt[2];
                    h = gt[3] * tanh_fast(cst[t]);
                }
                const __half h1 = __float2half_rn(h);
                const __half h2 = __float2half_rn((h - __half2float(h1)) * X2_SCALE);
                hv[t] = h; hh1[t] = h1; hh2[t] = h2; wr[t] = cval[t] && live;
                gather_rows8(h1, h2, lane, row1[t], row2[t]);
            }
            const long long e_cell = dbg_w ? clock64() : 0;
            if (s + 1 < maxlen) {
                const uint32_t boff = (uint32_t)((g * 2 + nxt) * B_BUF_B), bar = dstBar + (uint32_t)((g * 2 + nxt) * 2) * 8u;
#pragma unroll
                for (int t = 0; t < NT; ++t)
                    if (slot_live[t]) {                     // only the group's real line slots travel (and are counted by the receivers)
                        st_async_v4(dstB + boff + (uint32_t)(4 * t * 16), row1[t], bar);
                        st_async_v4(dstB + boff + (uint32_t)((LPW + 4 * t) * 16), row2[t], bar);
                    }
            }
            const long long e_sent = dbg_w ? clock64() : 0;
            // h to HBM and the gx prefetch come AFTER the hand-off: nothing on the critical path waits for them
#pragma unroll
            for (int t = 0; t < NT; ++t)
                if (wr[t]) {
                    if (p.out) p.out[ooff[t]] = hv[t];
                    if (p.out_hi) {
                        if (vec_planes) {
                            if (jq == 0) *reinterpret_cast<uint4 *>(p.out_hi + ooff[t]) = row1[t];
                            else if (jq == 1) *reinterpret_cast<uint4 *>(p.out_lo + ooff[t] - 1) = row2[t];
                        } else { p.out_hi[ooff[t]] = hh1[t]; p.out_lo[ooff[t]] = hh2[t]; }
                    }
                    ooff[t] += ostride;
                }
#pragma unroll
            for (int i = 0; i < LPW; ++i) {                                   // gx of the next step: a whole step to land
                gxn[i] = s + 1 < glen[i] ? __ldg(gptr[i]) : 0.f;
                gptr[i] += gstride;
            }
            if (dbg_w && s >= 100 && s < 104) {
                long long *d = p.dbgbuf + ((s - 100) * 2 + g) * 12;
                d[3] = e_top; d[4] = e_done; d[5] = e_ld; d[6] = e_act; d[7] = e_cell; d[8] = e_sent; d[9] = clock64();
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// Small-hidden variant (hid <= 32: the blla 2-D BiLSTM sweeps Lbx32 / Lby32): one CTA, no cluster.
//
//   All 4*hid <= 128 gate rows fit one CTA (row = 4*unit + gate), K = 32, so a CTA carries SNL = 64 sequences of one direction:
//   A: W1 | W2s planes, 128 rows x K32 -> 2 x 16 TMEM columns.    B: [h1 (64 lines) | h2s (64 lines)] x K32 in one SW128 tile
//   (128 rows x 128 B, k-chunks 0..3 used), double buffered.  4 MMAs per step: D1 (M128 N128 K16) x2, D2 (M128 N64 K16) x2.
//   The page's thousands of short rows/columns give 86..114 CTAs; a step is bounded by the 96 KB TMEM read-back and the
//   8192 SFU activations, not by the MMAs.  Same arithmetic as the clustered kernel (fp16 pairs, fp32 accumulate).
template <int NLS> struct SmallCfg {                    // NLS lines per CTA, EWS = NLS / 16 epilogue warps per TMEM lane quarter
    static constexpr int EWS = NLS / 16;
    static constexpr int THREADS = 32 + 4 * EWS * 32;
    static constexpr int B_BUF_B = 2 * NLS * 128;        // [h1 | h2s] rows x 128 B
    static constexpr int SG = NLS * 8 * 4, SH = NLS * 8;
    static constexpr int STG_BYTES = 4 * (SG + SH) * 4;
    static constexpr int INFO_BYTES = NLS * 8 + NLS * 8 + NLS * 4;   // per line: base pixel, first gx element (long long), length (int)
    static constexpr int SMEM_BYTES = 2 * B_BUF_B + STG_BYTES + INFO_BYTES + 128 + 1024;
    static constexpr int TM_A0 = 3 * NLS;                // D1 @0 (2 NLS cols), D2 @2 NLS (NLS cols); A: W1 @3 NLS (16 cols), W2s @3 NLS + 16
    static constexpr int TM_ALLOC = NLS == 64 ? 256 : NLS == 32 ? 128 : 128;
    static constexpr int MINB = NLS == 64 ? 1 : 2;
};
constexpr int S_LPW = 16;                                // lines per epilogue warp
constexpr int S_A_PLANE_ELEMS = 128 * 32;

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

template <int NLS>
__global__ void __launch_bounds__(SmallCfg<NLS>::THREADS, SmallCfg<NLS>::MINB) k_lstm_rec_tc_small(LstmTcParams p) {
    using Cfg = SmallCfg<NLS>;
    constexpr int SNL = NLS, S_B_BUF_B = Cfg::B_BUF_B, S_SG = Cfg::SG, S_SH = Cfg::SH, S_STG_BYTES = Cfg::STG_BYTES;
    constexpr int S_INFO_BYTES = Cfg::INFO_BYTES, S_TM_A0 = Cfg::TM_A0, EWS = Cfg::EWS, STHREADS = Cfg::THREADS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_align1024(smem_raw);
    uint8_t *sB = smem;
    float *stg = reinterpret_cast<float *>(sB + 2 * S_B_BUF_B);
    long long *lbase = reinterpret_cast<long long *>(sB + 2 * S_B_BUF_B + S_STG_BYTES);
    long long *lgoff = lbase + SNL;
    int *llen = reinterpret_cast<int *>(lgoff + SNL);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + 2 * S_B_BUF_B + S_STG_BYTES + S_INFO_BYTES);
    uint64_t *b_full = bars + 1 /* [2] */, *mma_done = bars + 3;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.x, dir = blockIdx.y;
    const int hid = p.hid, GC = p.dirs * 4 * hid, OC = p.dirs * hid;

    if (threadIdx.x == 0) {
        mbar_init(&b_full[0], 4 * EWS); mbar_init(&b_full[1], 4 * EWS); mbar_init(mma_done, 1);      // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 2 * S_B_BUF_B / 16; i += STHREADS) reinterpret_cast<uint4 *>(sB)[i] = make_uint4(0, 0, 0, 0);   // h_{-1} = 0, k >= 32 stays 0
    if ((int)threadIdx.x < SNL) {
        const int ql = chunk * SNL + threadIdx.x;
        const int qq = ql < p.nseq ? ql : 0;
        const long long lb = (long long)(qq / p.q2) * p.s_outer + (long long)(qq % p.q2) * p.s_inner;
        const int ln = ql < p.nseq ? min(max(p.lens ? p.lens[ql] : p.T, 0), p.T) : 0;
        lbase[threadIdx.x] = lb;
        llen[threadIdx.x] = ln;
        lgoff[threadIdx.x] = (lb + (long long)(dir ? max(ln - 1, 0) : 0) * p.step) * GC;      // gx element of the line's first time step
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TM_ALLOC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= 1 && warp <= 4) {
        const int m = 32 * (warp & 3) + lane;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(p.wpk) + ((size_t)dir * 2 * S_A_PLANE_ELEMS) / 2 + (size_t)m * 16;
#pragma unroll 1
        for (int pl = 0; pl < 2; ++pl) {
            uint32_t r[16];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)pl * (S_A_PLANE_ELEMS / 2) + i));
                r[i] = v.x; r[i + 1] = v.y; r[i + 2] = v.z; r[i + 3] = v.w;
            }
            tmem_st16(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(S_TM_A0 + pl * 16), r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    int maxlen = 0;
    for (int lb = 0; lb < SNL; ++lb) maxlen = max(maxlen, llen[lb]);

    if (warp == 0) {
        // ===================== MMA issuer =====================
        const uint32_t id1 = idesc_f16(0, 0, 128, 2 * SNL), id2 = idesc_f16(0, 0, 128, SNL);
        for (int s = 0; s < maxlen; ++s) {
            const int cur = s & 1;
            if (s > 0) mbar_wait(&b_full[cur], (uint32_t)(((s - 1) >> 1) & 1));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t b0 = smem_u32(sB + cur * S_B_BUF_B);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint64_t bd = umma_desc_sw128(b0) + (uint64_t)(2 * k);
                    const uint32_t a1 = tmem_base + (uint32_t)(S_TM_A0 + k * 8);
                    umma_f16_ts(tmem_base, a1, bd, id1, (uint32_t)k);                       // [W1 h1 | W1 h2s]
                    umma_f16_ts(tmem_base + (uint32_t)(2 * SNL), a1 + 16u, bd, id2, (uint32_t)k);   // W2s h1 (first 64 rows of B)
                }
                umma_commit(mma_done);
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue warps 1..16: quarter q = warp & 3, sub-warp sw = (warp-1) >> 2 =====================
        const int q = warp & 3;
        const int sw = (warp - 1) >> 2;
        const int jq = lane >> 2, g = lane & 3;
        const int u = 8 * q + jq;
        const bool uvalid = u < hid;
        float *sg = stg + q * (S_SG + S_SH), *sh = sg + S_SG;
        const int tq = sw * 32 + lane;
        // the four cells this thread updates: line cl, units cu0 .. cu0+3
        const int cl = tq >> 1, cj0 = (tq & 1) * 4, cu0 = 8 * q + cj0;
        const int cql = chunk * SNL + cl;
        const int clen = llen[cl];
        const long long cbase = lbase[cl];
        float cst[4] = {0.f, 0.f, 0.f, 0.f};
        if (cql < p.nseq)
            for (int tt = clen; tt < p.T; ++tt)
                for (int e = 0; e < 4; ++e)
                    if (cu0 + e < hid) {
                        const size_t o = (size_t)(cbase + (long long)tt * p.step) * OC + dir * hid + cu0 + e;
                        if (p.out) p.out[o] = 0.f;
                        if (p.out_hi) { p.out_hi[o] = __float2half_rn(0.f); p.out_lo[o] = __float2half_rn(0.f); }
                    }
        long long ooff = (cbase + (long long)(dir ? max(clen - 1, 0) : 0) * p.step) * OC + dir * hid + cu0;
        const long long ostride = (long long)(dir ? -1 : 1) * p.step * OC;
        const bool ovec = (hid & 3) == 0 && cu0 + 3 < hid;
        const float *gx0 = p.gx + (size_t)dir * 4 * hid + (size_t)(uvalid ? u : 0) * 4 + g;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(S_LPW * sw);
        const float act_k = g == 2 ? 2.f : 1.f;

        // gx of this thread's gate row for its 16 lines, fetched one time step ahead (the loads have a whole step to land)
        const long long gstride = (long long)(dir ? -1 : 1) * p.step * GC;
        float gxn[S_LPW];
#pragma unroll
        for (int i = 0; i < S_LPW; ++i) {
            const int l = S_LPW * sw + i;
            gxn[i] = (uvalid && 0 < llen[l]) ? __ldg(gx0 + lgoff[l]) : 0.f;
        }
        for (int s = 0; s < maxlen; ++s) {
            const int nxt = (s + 1) & 1;
            float gxv[S_LPW];
            {
                const float *gs = gx0 + (long long)(s + 1) * gstride;
#pragma unroll
                for (int i = 0; i < S_LPW; ++i) {
                    const int l = S_LPW * sw + i;
                    gxv[i] = gxn[i];
                    gxn[i] = (uvalid && s + 1 < llen[l]) ? __ldg(gs + lgoff[l]) : 0.f;
                }
            }
            mbar_wait(mma_done, (uint32_t)(s & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t m0[S_LPW], c0[S_LPW], c1[S_LPW];
            tmem_ld16_nowait(lane_base, m0);
            tmem_ld16_nowait(lane_base + SNL, c0);
            tmem_ld16_nowait(lane_base + 2 * SNL, c1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#pragma unroll
            for (int i = 0; i < S_LPW; ++i) {
                const float corr = __uint_as_float(c0[i]) + __uint_as_float(c1[i]);
                const float pre = (__uint_as_float(m0[i]) + corr * (1.f / X2_SCALE)) + gxv[i];
                sg[((S_LPW * sw + i) * 8 + jq) * 4 + g] = fmaf(sigmoid_fast(pre * act_k), act_k, 1.f - act_k);
            }
            named_bar(1 + q, 32 * EWS);
            {
                float hv[4];
                const bool live = cql < p.nseq && s < clen;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 gt = *reinterpret_cast<const float4 *>(&sg[(cl * 8 + cj0 + e) * 4]);      // i, f, g, o
                    hv[e] = 0.f;
                    if (live && cu0 + e < hid) {
                        cst[e] = gt.y * cst[e] + gt.x * gt.z;
                        hv[e] = gt.w * tanh_fast(cst[e]);
                    }
                }
                if (live) {
                    if (p.out) {
                        if (ovec) *reinterpret_cast<float4 *>(p.out + ooff) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                        else
                            for (int e = 0; e < 4; ++e) if (cu0 + e < hid) p.out[ooff + e] = hv[e];
                    }
                    if (p.out_hi) {
                        __half a[4], b[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) { a[e] = __float2half_rn(hv[e]); b[e] = __float2half_rn((hv[e] - __half2float(a[e])) * X2_SCALE); }
                        if (ovec) {
                            *reinterpret_cast<uint2 *>(p.out_hi + ooff) = make_uint2(pack_h2(a[0], a[1]), pack_h2(a[2], a[3]));
                            *reinterpret_cast<uint2 *>(p.out_lo + ooff) = make_uint2(pack_h2(b[0], b[1]), pack_h2(b[2], b[3]));
                        } else {
                            for (int e = 0; e < 4; ++e) if (cu0 + e < hid) { p.out_hi[ooff + e] = a[e]; p.out_lo[ooff + e] = b[e]; }
                        }
                    }
                    ooff += ostride;
                }
                *reinterpret_cast<float4 *>(&sh[cl * 8 + cj0]) = make_float4(hv[0], hv[1], hv[2], hv[3]);
            }
            named_bar(1 + q, 32 * EWS);
            if (s + 1 < maxlen) {
                // chunk = 8 units of one line in one fp16 plane: tile row = plane*64 + line, 16-byte k-chunk q
                const int plane = tq / SNL, line = tq % SNL;
                const float4 x0 = *reinterpret_cast<const float4 *>(&sh[line * 8]), x1 = *reinterpret_cast<const float4 *>(&sh[line * 8 + 4]);
                const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    uint32_t two[2];
#pragma unroll
                    for (int f = 0; f < 2; ++f) {
                        const float x = xs[2 * e + f];
                        const __half h1 = __float2half_rn(x);
                        const __half h2 = __float2half_rn((x - __half2float(h1)) * X2_SCALE);
                        two[f] = (uint32_t)__half_as_ushort(plane == 0 ? h1 : h2);
                    }
                    pk[e] = two[0] | (two[1] << 16);
                }
                const int row = plane * SNL + line;
                const uint32_t off = (uint32_t)(nxt * S_B_BUF_B + row * 128 + ((q ^ (row & 7)) << 4));
                // plain shared store + proxy fence + one mbarrier arrival per warp (st.async faults with "illegal instruction"
                // in a launch without a cluster dimension, measured)
                *reinterpret_cast<uint4 *>(sB + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&b_full[nxt]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TM_ALLOC) : "memory");
    }
}

}  // namespace ltc
}  // namespace kb
