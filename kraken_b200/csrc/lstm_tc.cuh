// lstm_tc.cuh - LSTM recurrence on tcgen05 tensor cores for hidden sizes 129..256 (sm_100a).
// EXPERIMENTAL, opt-in with KB_LSTM_TC=1: bit-for-bit label parity like the default kernel (tests/test_gpu_parity.py::
// test_tensor_core_recurrence) but measured 2.4x SLOWER on cfg2 (1.63 ms vs 0.69 ms per 64 x 200 steps): a step is 96
// dependent M128xN16xK16 MMAs on two accumulators (accumulator-latency bound, ~12k cycles) followed by the serialised
// pointwise + all-to-all exchange.  Kept as the starting point for the N=48 merged-operand variant described in DESIGN.md.
//
// Same contract as k_lstm_rec (kernels.cuh): per-pixel gate pre-activations gx in, hidden states out, packed-sequence
// semantics, one direction per blockIdx.y.  What changes is where  W_hh . h_{t-1}  is computed:
//
//   cluster of 8 CTAs = 16 sequences of one direction for all time steps; CTA r owns hidden units [32r, 32r+32)
//   (unit slots are padded to 32 per CTA, so any hid <= 256 works), i.e. 128 gate rows, ordered unit-major
//   (row = 4*unit + gate) so that the 4 gates of a unit sit in 4 adjacent TMEM lanes.
//
//   A = W_hh slice [128 rows][K = 256], split into THREE bf16 planes (w = w1 + w2 + w3, 24 significand bits), resident in
//       shared memory for the whole kernel in the UMMA K-major 128B-swizzle layout (pre-swizzled on the host,
//       fetched with 12 cp.async.bulk copies): 192 KB.
//   B = h_{t-1}^T [16 lines][K = 256], three bf16 planes as well (24 KB), written REMOTELY: every CTA converts the h_t
//       of its 32 units to bf16x3 and st.async.v4's the 16-byte chunks into all 8 CTAs' B tiles; the destination's
//       mbarrier counts the bytes (no cluster barrier, no fence).
//   D = 96 x tcgen05.mma.kind::f16 (M128 x N16 x K16) per step into two TMEM accumulators:
//       main = w1*h1 (16 accumulations - keeps the tensor core's round-toward-zero chain short),
//       corr = w1*h2 + w2*h1 + w2*h2 + w1*h3 + w3*h1   (dropped terms <= 2^-24).
//   epilogue: tcgen05.ld -> + gx -> sigmoid/tanh (accurate expf/tanhf) -> gates regrouped through shared memory ->
//       fp32 cell update (4 cells per thread) -> h_t to HBM and to every CTA's B operand.
//   B is single-buffered (shared memory is full), so a CTA signals "my MMAs of this step have retired" to all CTAs
//   (remote mbarrier arrive) and senders wait for 8 such signals before overwriting B.
#pragma once
#include <cuda_bf16.h>

#include "gemm_tc.cuh"

namespace kb {
namespace ltc {

using namespace kb::tc;

constexpr int NL = 16;                                   // lines per cluster (= MMA N)
constexpr int LCS = 8;                                   // cluster size
constexpr int A_TILE_B = 128 * 128;                      // one (split, k-atom) tile of A: 128 rows x 128 B
constexpr int A_BYTES = 3 * 4 * A_TILE_B;                // 196608
constexpr int B_TILE_B = NL * 128;                       // one (split, k-atom) tile of B: 16 rows x 128 B
constexpr int B_BYTES = 3 * 4 * B_TILE_B;                // 24576
constexpr int STG_BYTES = 4 * NL * 8 * 4 * 4;            // per warp [16 lines][8 units][4 gates] fp32 = 2 KB -> 8 KB
constexpr int LSMEM_BYTES = A_BYTES + B_BYTES + STG_BYTES + 128 + 1024;
constexpr int LTHREADS = 160;                            // warp 0: MMA issue / TMEM; warps 1..4: epilogue

struct LstmTcParams {
    const float *gx; const __nv_bfloat16 *wpk; float *out; const int *lens;
    int nseq, T, hid, dirs, U;
    int q2; long long s_outer, s_inner, step;
};

__device__ __forceinline__ uint32_t idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mapa32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint4 v, uint32_t rmbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void remote_arrive(uint32_t rmbar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rmbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {      // acquire at cluster scope (remote arrivals)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "CW_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra CD_%=;\n\t"
        "bra CW_%=;\n\t"
        "CD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// x -> three bf16 terms, x == b1 + b2 + b3 up to 2^-24 |x|
__device__ __forceinline__ void split3(float x, __nv_bfloat16 &b1, __nv_bfloat16 &b2, __nv_bfloat16 &b3) {
    b1 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(b1);
    b2 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(b2);
    b3 = __float2bfloat16_rn(r2);
}

__global__ void __launch_bounds__(LTHREADS, 1) k_lstm_rec_tc(LstmTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sA = smem, *sB = smem + A_BYTES;
    float *stg = reinterpret_cast<float *>(sB + B_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + B_BYTES + STG_BYTES);
    uint64_t *a_full = bars, *b_full = bars + 1, *b_free = bars + 2, *mma_done = bars + 3;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int chunk = blockIdx.x / LCS, dir = blockIdx.y;
    const int hid = p.hid, GC = p.dirs * 4 * hid, OC = p.dirs * hid;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1); mbar_init(b_full, 1); mbar_init(b_free, LCS); mbar_init(mma_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(a_full, A_BYTES);
        const uint8_t *src = reinterpret_cast<const uint8_t *>(p.wpk) + ((size_t)dir * LCS + rank) * A_BYTES;
        for (int i = 0; i < 12; ++i) bulk_g2s(sA + i * A_TILE_B, src + (size_t)i * A_TILE_B, A_TILE_B, a_full);
        mbar_expect_tx(b_full, B_BYTES);                 // first fill: the h_0 every CTA sends at the end of step 0
    }
    for (int i = threadIdx.x; i < B_BYTES / 16; i += LTHREADS) reinterpret_cast<uint4 *>(sB)[i] = make_uint4(0, 0, 0, 0);   // h_{-1} = 0
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // generic-proxy zero fill of B must be visible to the async proxy (UMMA reads)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");

    // sequence lengths of the 16 lines of this cluster (uniform across the cluster)
    int maxlen = 0;
    for (int lb = 0; lb < NL; ++lb) {
        const int q = chunk * NL + lb;
        if (q < p.nseq) maxlen = max(maxlen, p.lens ? min(max(p.lens[q], 0), p.T) : p.T);
    }

    if (warp == 0) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = idesc_bf16(128, NL);
        const uint32_t d_main = tmem_base, d_corr = tmem_base + NL;
        mbar_wait(a_full, 0);
        for (int s = 0; s < maxlen; ++s) {
            if (s > 0) mbar_wait_cluster(b_full, (uint32_t)((s - 1) & 1));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // st.async writes -> visible to the UMMA (async proxy) reads
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                if (s > 0 && s + 1 < maxlen) mbar_expect_tx(b_full, B_BYTES);      // next fill (sent at the end of this step)
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
                // (A split, B split) pairs: main first, then the five correction products
                const int pa[6] = {0, 0, 1, 1, 0, 2}, pb[6] = {0, 1, 0, 1, 2, 0};
#pragma unroll
                for (int pr = 0; pr < 6; ++pr) {
                    const uint32_t d = pr == 0 ? d_main : d_corr;
#pragma unroll
                    for (int ka = 0; ka < 4; ++ka) {
                        const uint64_t ad = umma_desc_sw128(a0 + (uint32_t)((pa[pr] * 4 + ka) * A_TILE_B));
                        const uint64_t bd = umma_desc_sw128(b0 + (uint32_t)((pb[pr] * 4 + ka) * B_TILE_B));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (pr <= 1 && ka == 0 && k == 0) ? 0u : 1u);
                    }
                }
                umma_commit(mma_done);
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue warps 1..4 =====================
        const int q = warp & 3;                            // TMEM lane quarter: lanes 32q .. 32q+31 = unit slots 8q .. 8q+7
        const int jq = lane >> 2, g = lane & 3;            // unit slot within the warp, gate (i,f,g,o) of this lane's TMEM row
        const int slot = 8 * q + jq;                       // unit slot in the CTA (0..31)
        const int u = (int)rank * p.U + slot;              // real hidden unit
        const bool uvalid = slot < p.U && u < hid;
        float *wstg = stg + q * (NL * 8 * 4);              // this warp's [line][unit][gate] staging
        // the 4 cells this thread updates: unit jq, lines 4b + g  (after the regroup lane g of a unit owns lines = g mod 4)
        int len4[4]; long long base4[4]; float cst[4]; bool lv[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int ql = chunk * NL + 4 * b + g;
            lv[b] = ql < p.nseq;
            const int l = lv[b] ? (p.lens ? p.lens[ql] : p.T) : 0;
            len4[b] = min(max(l, 0), p.T);
            const int qq = lv[b] ? ql : 0;
            base4[b] = (long long)(qq / p.q2) * p.s_outer + (long long)(qq % p.q2) * p.s_inner;
            cst[b] = 0.f;
        }
        // zero the padded tails (pad_packed_sequence) of this thread's cells
        if (uvalid)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (lv[b]) for (int t = len4[b]; t < p.T; ++t) p.out[(size_t)(base4[b] + (long long)t * p.step) * OC + dir * hid + u] = 0.f;
        // lengths/bases of all 16 lines for the gx loads of this lane's TMEM row (gate g of unit jq, every line)
        uint32_t rB[LCS], rFull[LCS], rFree[LCS];
#pragma unroll
        for (int r = 0; r < LCS; ++r) {
            rB[r] = mapa32(smem_u32(sB), (uint32_t)r); rFull[r] = mapa32(smem_u32(b_full), (uint32_t)r); rFree[r] = mapa32(smem_u32(b_free), (uint32_t)r);
        }
        for (int s = 0; s < maxlen; ++s) {
            // gate pre-activations of x for this row: line i at its own time index
            float gxv[NL];
#pragma unroll
            for (int i = 0; i < NL; ++i) {
                const int ql = chunk * NL + i;
                float v = 0.f;
                if (uvalid && ql < p.nseq) {
                    const int l = p.lens ? min(max(p.lens[ql], 0), p.T) : p.T;
                    if (s < l) {
                        const int t = dir ? l - 1 - s : s;
                        const long long bs = (long long)(ql / p.q2) * p.s_outer + (long long)(ql % p.q2) * p.s_inner;
                        v = __ldg(p.gx + (size_t)(bs + (long long)t * p.step) * GC + (size_t)dir * 4 * hid + (size_t)u * 4 + g);
                    }
                }
                gxv[i] = v;
            }
            mbar_wait(mma_done, (uint32_t)(s & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // this CTA's MMAs have retired: its B operand may be overwritten -> tell every CTA of the cluster
            if (warp == 1 && lane < LCS && s + 1 < maxlen) remote_arrive(rFree[lane]);
            float dm[NL], dc[NL];
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
            tmem_ld16(lane_base, dm);
            tmem_ld16(lane_base + NL, dc);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#pragma unroll
            for (int i = 0; i < NL; ++i) {
                const float pre = (dm[i] + dc[i]) + gxv[i];
                const float a = g == 2 ? tanhf(pre) : sigmoidf_acc(pre);
                wstg[(i * 8 + jq) * 4 + g] = a;
            }
            __syncwarp();
            float hv[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int line = 4 * b + g;
                const float4 gt = *reinterpret_cast<const float4 *>(&wstg[(line * 8 + jq) * 4]);     // i, f, g, o
                const bool act = uvalid && lv[b] && s < len4[b];
                if (act) {
                    cst[b] = gt.y * cst[b] + gt.x * gt.z;
                    hv[b] = gt.w * tanhf(cst[b]);
                    const int t = dir ? len4[b] - 1 - s : s;
                    p.out[(size_t)(base4[b] + (long long)t * p.step) * OC + dir * hid + u] = hv[b];
                } else hv[b] = 0.f;                        // finished / padding cells feed zeros (their h is never used again)
            }
            __syncwarp();
            if (s + 1 < maxlen) {
                // h_t of the warp's 8 unit slots x 16 lines -> staging as [line][8 units] fp32, then 16-byte bf16 chunks
                float *hst = wstg;                         // reuse: [16][8]
#pragma unroll
                for (int b = 0; b < 4; ++b) hst[(4 * b + g) * 8 + jq] = hv[b];
                __syncwarp();
                const int line = lane & 15, part = lane >> 4;      // lanes 0..15: splits 0,1 of line; lanes 16..31: split 2
                const float4 x0 = *reinterpret_cast<const float4 *>(&hst[line * 8]), x1 = *reinterpret_cast<const float4 *>(&hst[line * 8 + 4]);
                const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                __nv_bfloat16 s1[8], s2[8], s3[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split3(xs[e], s1[e], s2[e], s3[e]);
                auto pack = [](const __nv_bfloat16 *v) {
                    uint4 o;
                    o.x = (uint32_t)__bfloat16_as_ushort(v[0]) | ((uint32_t)__bfloat16_as_ushort(v[1]) << 16);
                    o.y = (uint32_t)__bfloat16_as_ushort(v[2]) | ((uint32_t)__bfloat16_as_ushort(v[3]) << 16);
                    o.z = (uint32_t)__bfloat16_as_ushort(v[4]) | ((uint32_t)__bfloat16_as_ushort(v[5]) << 16);
                    o.w = (uint32_t)__bfloat16_as_ushort(v[6]) | ((uint32_t)__bfloat16_as_ushort(v[7]) << 16);
                    return o;
                };
                // position of this warp's 8 unit slots in K: k0 = 32*rank + 8*q -> k-atom ka, 16-byte chunk c, swizzled with the row
                const int k0 = (int)rank * 32 + 8 * q, ka = k0 >> 6, c = (k0 & 63) >> 3;
                const uint32_t off_in_tile = (uint32_t)(line * 128 + ((c ^ (line & 7)) << 4));
                // every CTA must have finished reading its B operand for this step
                mbar_wait_cluster(b_free, (uint32_t)(s & 1));
                if (part == 0) {
                    const uint4 c1 = pack(s1), c2 = pack(s2);
#pragma unroll
                    for (int r = 0; r < LCS; ++r) {
                        st_async_v4(rB[r] + (uint32_t)((0 * 4 + ka) * B_TILE_B) + off_in_tile, c1, rFull[r]);
                        st_async_v4(rB[r] + (uint32_t)((1 * 4 + ka) * B_TILE_B) + off_in_tile, c2, rFull[r]);
                    }
                } else {
                    const uint4 c3 = pack(s3);
#pragma unroll
                    for (int r = 0; r < LCS; ++r) st_async_v4(rB[r] + (uint32_t)((2 * 4 + ka) * B_TILE_B) + off_in_tile, c3, rFull[r]);
                }
                __syncwarp();
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32) : "memory");
    }
}

}  // namespace ltc
}  // namespace kb
