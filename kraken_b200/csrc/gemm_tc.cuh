// gemm_tc.cuh - tcgen05 / TMEM / TMA GEMM with fp32-grade accuracy ("3xTF32") for sm_100a.
//
//   C[M][N] = A[M][K] * B[N][K]^T + bias[N]          (A: activations, pixel-major; B: weights, K-major)
//
// Used for the LSTM input projection (K = 768 -> N = 2048 on cfg2, 50% of the path's FLOPs), the output Linear and
// 1x1 convolutions.  CTC label sequences must match the fp32 reference bit for bit, and a single TF32/BF16 pass
// flips arg-maxes, so every fp32 operand is split once into two TF32-exact terms
//       x = hi + lo,   hi = rna_tf32(x),   lo = x - hi   (exact; |lo| <= 2^-11 |x|)
// and three tensor-core products are accumulated in fp32 in TMEM:
//       main = a_hi*b_hi            corr = a_lo*b_hi + a_hi*b_lo            (dropped term a_lo*b_lo ~ 2^-22)
// The tensor core adds into its accumulator with round-toward-zero, a bias that grows linearly with the number of
// accumulations into a LARGE accumulator (measured: 4.2e-6 relative at K = 768 with one accumulator vs 7.6e-7 for
// fp32 FFMA).  Keeping the small correction products in their own accumulator (columns 256..511) cuts the chain
// on the big one to K/8 and makes the correction's own rounding negligible; the epilogue adds the two in fp32 (RN).
// The split planes live in HBM (weights: once at finalize; activations: k_split_tf32 below) so the kernel's
// shared-memory bandwidth is spent on TMA fills and UMMA operand reads only.
//
// Kernel anatomy (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer : 4 x cp.async.bulk.tensor.2d (A_hi, A_lo, B_hi, B_lo; 128B swizzle) per k-block
//   warp 1      MMA issuer   : tcgen05.mma.cta_group::1.kind::tf32, M128 x N256 x K8, 12 per k-block of 32 floats;
//               also owns TMEM alloc/dealloc (512 columns = main + correction accumulator of 256 each)
//   warps 2..5  epilogue     : tcgen05.ld 32x32b.x32 -> + bias -> st.global (one accumulator row per thread)
// Pipelines: smem full/empty mbarriers (2 stages x 96 KB), TMEM full/empty mbarrier pair (one accumulator set).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace kb {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 32, STAGES = 2;
constexpr int A_TILE = BM * BK * 4, B_TILE = BN * BK * 4;                 // bytes
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;                      // 96 KB
constexpr int EPI_LD = 36;                                                // padded row of the per-warp 32x32 staging tile
constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;                            // 4 epilogue warps
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_BYTES;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// 1024-byte alignment of the dynamic shared window WITHOUT an integer round trip: a uintptr_t cast loses the address space and
// every later access compiles to generic LD/ST with 64-bit address arithmetic (seen in the ncu source page, r01c)
__device__ __forceinline__ uint8_t *smem_align1024(uint8_t *p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// K-major operand tile [rows][32 fp32] written by TMA with CU_TENSOR_MAP_SWIZZLE_128B: rows are 128 B apart,
// 8-row groups 1024 B apart (SBO), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, 16-byte units
    d |= (uint64_t)1 << 16;                           // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // SBO
    d |= (uint64_t)1 << 46;                           // descriptor version
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct GemmTcParams {
    float *c; const float *bias;
    int M, N, K, ldc, act;
};

__global__ void __launch_bounds__(THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
          const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, GemmTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_align1024(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + STAGES, *tfull = bars + 2 * STAGES, *tempty = bars + 2 * STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);
    float *epi_stage = reinterpret_cast<float *>(smem + STAGES * STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
    const int ntiles = tiles_m * tiles_n;
    const int nkb = (p.K + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&tfull[0], 1); mbar_init(&tempty[0], 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                // n-tiles innermost: consecutive CTAs share the same A rows (L2 reuse), weights stay L2 resident
                const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t *st = smem + stage * STAGE_BYTES;
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    tma_load_2d(st, &tm_a_hi, &full[stage], kb * BK, m0);
                    tma_load_2d(st + A_TILE, &tm_a_lo, &full[stage], kb * BK, m0);
                    tma_load_2d(st + 2 * A_TILE, &tm_b_hi, &full[stage], kb * BK, n0);
                    tma_load_2d(st + 2 * A_TILE + B_TILE, &tm_b_lo, &full[stage], kb * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = idesc_tf32(BM, BN);
        int stage = 0; uint32_t phase = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            mbar_wait(&tempty[0], acc_phase ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_main = tmem_base, d_corr = tmem_base + (uint32_t)BN;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + A_TILE);
                    const uint64_t b_hi = umma_desc_sw128(sa + 2 * A_TILE), b_lo = umma_desc_sw128(sa + 2 * A_TILE + B_TILE);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);      // 32 bytes per K=8 step inside the 128B swizzle atom
                        umma_tf32(d_corr, a_lo + adv, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
                        umma_tf32(d_corr, a_hi + adv, b_lo + adv, idesc, 1u);
                        umma_tf32(d_main, a_hi + adv, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);                               // frees the smem stage when these MMAs retire
                    if (kb == nkb - 1) umma_commit(&tfull[0]);                // accumulators complete -> epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            acc_phase ^= 1;
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                                              // TMEM lane quarter this warp may access
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
            mbar_wait(&tfull[0], acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float *stg = epi_stage + (warp - 2) * 32 * EPI_LD;
            const bool vec = (p.ldc & 3) == 0 && (p.N & 3) == 0;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                if (n0 + c0 >= p.N) break;                                    // warp-uniform
                float v[32], cr[32];
                const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
                tmem_ld32(lane_base + (uint32_t)c0, v);
                tmem_ld32(lane_base + (uint32_t)(BN + c0), cr);
                // row = lane: main + correction (+ bias, activation), staged so that the global stores below are
                // row-contiguous (each store instruction writes 4 rows x 128 B instead of 32 rows x 16 B)
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int n = n0 + c0 + j;
                    float4 o;
                    o.x = v[j] + cr[j]; o.y = v[j + 1] + cr[j + 1]; o.z = v[j + 2] + cr[j + 2]; o.w = v[j + 3] + cr[j + 3];
                    if (p.bias) {
                        o.x += (n + 0 < p.N) ? __ldg(p.bias + n + 0) : 0.f; o.y += (n + 1 < p.N) ? __ldg(p.bias + n + 1) : 0.f;
                        o.z += (n + 2 < p.N) ? __ldg(p.bias + n + 2) : 0.f; o.w += (n + 3 < p.N) ? __ldg(p.bias + n + 3) : 0.f;
                    }
                    o.x = act_apply(o.x, p.act); o.y = act_apply(o.y, p.act); o.z = act_apply(o.z, p.act); o.w = act_apply(o.w, p.act);
                    *reinterpret_cast<float4 *>(stg + lane * EPI_LD + j) = o;
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = i * 4 + (lane >> 3), c4 = (lane & 7) * 4;
                    const int m = m0 + q * 32 + r, n = n0 + c0 + c4;
                    if (m < p.M && n < p.N) {
                        const float4 o = *reinterpret_cast<const float4 *>(stg + r * EPI_LD + c4);
                        float *dst = p.c + (size_t)m * p.ldc + n;
                        if (vec) *reinterpret_cast<float4 *>(dst) = o;
                        else { dst[0] = o.x; if (n + 1 < p.N) dst[1] = o.y; if (n + 2 < p.N) dst[2] = o.z; if (n + 3 < p.N) dst[3] = o.w; }
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[0]);
            acc_phase ^= 1;
        }
    }
    // ===================== teardown =====================
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// x -> (hi, lo): hi = round-to-nearest TF32 (low 13 mantissa bits zero), lo = x - hi (exact in fp32)
__global__ void k_split_tf32(const float *__restrict__ x, float *__restrict__ hi, float *__restrict__ lo, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(x) + i);
        float4 h, l;
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); h.x = __uint_as_float(t); l.x = v.x - h.x;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); h.y = __uint_as_float(t); l.y = v.y - h.y;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); h.z = __uint_as_float(t); l.z = v.z - h.z;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); h.w = __uint_as_float(t); l.w = v.w - h.w;
        reinterpret_cast<float4 *>(hi)[i] = h;
        reinterpret_cast<float4 *>(lo)[i] = l;
    }
}

// ---- host side: tensor maps -------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr; cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// row-major fp32 matrix [rows][K] (K contiguous), box = [box_rows][32], 128B swizzle, OOB -> 0
inline bool make_map_2d(CUtensorMap *map, const float *base, uint64_t rows, uint64_t K, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {K, rows};
    cuuint64_t strides[1] = {K * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc
}  // namespace kb
