// conv_tc.cuh - tcgen05 implicit-GEMM convolution (Cin a multiple of 32, stride 1, no dilation) with fused epilogue for sm_100a.
//
//   ActConv2D -> [Dropout] -> [MaxPool 2x2/2] -> [Dropout] -> [Reshape S1(1x0)1,3] -> (fp16 planes for the next tensor-core layer)
//   kraken/lib/vgsl/layers.py:842-852, 381-388, 313-335
//
// Work item = one image n, one PAIR of output rows (h0, h0+1), 128 output columns, one tile of <= 128 output channels.  In
// NHWC fp16 planes a 32-channel chunk of one pixel is exactly one 64-byte swizzle row, so the im2col matrix never exists (for
// Cin > 32 the item loops over the 32-channel chunks, re-using the row buffers and accumulating into the same TMEM columns):
//   * TMA (4-D tensor map C,W,H,N; box 32 x (128+kw-1) x 1 x 1, 64B swizzle) brings the kh+1 input rows a row pair
//     needs into shared memory ONCE; out-of-bounds coordinates are zero-filled by the TMA unit = the conv's zero padding
//   * the A operand of tap (ky, kx) for output row r is the row buffer (ky + r) with its UMMA descriptor start address
//     advanced by kx pixel rows (64 B each).  Measured on B200 (with 128-byte rows in round 1a, tools/gpu_fuse_debug.py): the
//     swizzle XOR is taken from the ABSOLUTE shared-memory address bits, exactly as TMA wrote it, so a descriptor may start
//     at any pixel row of an aligned buffer with base_offset = 0 (setting base_offset = kx gives garbage)
//   * B operand = the tap's [Cout][32] weight slice, streamed through a 4-stage TMA ring
//   * split fp16 operands as in gemm_tc.cuh: per tap and K=16 step  corr += a2s*b1 + a1*b2s ; main += a1*b1, for both
//     output rows -> 4 TMEM accumulators of Cout columns (two sets when 8*Cout <= 512, so the epilogue overlaps the MMAs)
//   * epilogue: one thread = one output column of the tile and holds BOTH rows -> the vertical half of the 2x2 max-pool is
//     a register max, the horizontal half one shuffle; bias/activation; the store is strided so that the `S` fold
//     (h into the feature axis) costs nothing; optional fp16 planes for the consumer.
#pragma once
#include "gemm_tc.cuh"

namespace kb {
namespace ctc {

using namespace kb::tc;

constexpr int TW = 128;                  // output columns per work item
constexpr int MAX_STB = 8;               // weight-tile ring stages (as many as fit next to the row buffers, at least 4)
constexpr int PIX_B = 64;                // bytes of one pixel row of a plane: 32 channels x fp16
constexpr int CEPI_WARPS = 8;             // two per TMEM lane quarter: even / odd 32-channel chunks
constexpr int CTHREADS = 64 + CEPI_WARPS * 32;
constexpr int MAX_ROWS = 8;              // kh + 1 <= 8

struct ConvTcParams {
    const float *bias; float *y; __half *y_hi; __half *y_lo; int *flag;
    int N, Ho, Wo, Cout, kh, kw, py, px, act, pool;
    int items_h, items_w, items_c;       // row pairs, column segments, output-channel tiles
    int NC, CT, nstb;                    // 32-channel input chunks, output channels per tile, weight ring stages
    long long sN, sH, sW;                // output strides (floats) of (n, out row, out col); channel stride 1
    int out_h, out_w;                    // valid output extent (pooled when pool)
    int a_row_bytes;                     // bytes of one plane of one input-row buffer (multiple of 1024)
    int acc_sets;
    int a_sets;                          // 2: the input rows of the NEXT (item, chunk) load while the current one is multiplied
    int w_res;                           // 1: all kh*kw*NC weight tiles stay resident in shared memory (loaded once per CTA); nstb = their number
};

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__global__ void __launch_bounds__(CTHREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
          const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo, ConvTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_align1024(smem_raw);
    const int R = p.kh + 1;
    const int a_plane = p.a_row_bytes, a_row = 2 * a_plane;              // hi | lo
    const int CT = p.CT, NSTB = p.nstb;
    const int b_plane = CT * PIX_B, b_stage = 2 * b_plane;
    const int ASETS = p.a_sets;
    uint8_t *a_base = smem, *b_base = smem + ASETS * R * a_row;
    uint64_t *bars = reinterpret_cast<uint64_t *>(b_base + NSTB * b_stage);
    uint64_t *full_a = bars /* [2][MAX_ROWS] */, *empty_a = bars + 2 * MAX_ROWS, *full_b = bars + 4 * MAX_ROWS, *empty_b = full_b + MAX_STB;
    uint64_t *tfull = empty_b + MAX_STB, *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nitems = p.N * p.items_h * p.items_w * p.items_c;
    const uint32_t a_tx = (uint32_t)(2 * (TW + p.kw - 1) * PIX_B), b_tx = (uint32_t)b_stage;

    if (threadIdx.x == 0) {
        for (int r = 0; r < 2 * MAX_ROWS; ++r) { mbar_init(&full_a[r], 1); mbar_init(&empty_a[r], 1); }
        for (int s = 0; s < (p.w_res ? 1 : NSTB); ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], CEPI_WARPS); }
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
            uint32_t a_phase = 0; int aset = 0; int bs = 0; uint32_t b_phase = 0;
            if (p.w_res && (int)blockIdx.x < nitems) {                      // the whole filter bank once (items_c == 1: tile = all output channels)
                const int ntaps = p.kh * p.kw * p.NC;
                mbar_expect_tx(&full_b[0], (uint32_t)ntaps * b_tx);
                for (int ti = 0; ti < ntaps; ++ti) {
                    tma_load_2d(b_base + ti * b_stage, &tm_w_hi, &full_b[0], 0, ti * p.Cout);
                    tma_load_2d(b_base + ti * b_stage + b_plane, &tm_w_lo, &full_b[0], 0, ti * p.Cout);
                }
            }
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int ct = item % p.items_c; int rest = item / p.items_c;
                const int ws = rest % p.items_w, hp = (rest / p.items_w) % p.items_h, n = rest / (p.items_w * p.items_h);
                const int h0 = 2 * hp - p.py, w0 = ws * TW - p.px;
                for (int cc = 0; cc < p.NC; ++cc) {
                    int r_loaded = 0;
                    uint8_t *aset_base = a_base + aset * R * a_row;
                    uint64_t *fa = full_a + aset * MAX_ROWS, *ea = empty_a + aset * MAX_ROWS;
                    // interleave: input row r is needed from tap row ky = r - 1 on; weight tiles in (ky, kx) order
                    for (int ky = 0; ky < p.kh; ++ky) {
                        for (; r_loaded <= ky + 1; ++r_loaded) {
                            mbar_wait(&ea[r_loaded], a_phase ^ 1);
                            mbar_expect_tx(&fa[r_loaded], a_tx);
                            tma_load_4d(aset_base + r_loaded * a_row, &tm_x_hi, &fa[r_loaded], cc * 32, w0, h0 + r_loaded, n);
                            tma_load_4d(aset_base + r_loaded * a_row + a_plane, &tm_x_lo, &fa[r_loaded], cc * 32, w0, h0 + r_loaded, n);
                        }
                        for (int kx = 0; kx < p.kw && !p.w_res; ++kx) {
                            mbar_wait(&empty_b[bs], b_phase ^ 1);
                            mbar_expect_tx(&full_b[bs], b_tx);
                            const int wrow = ((ky * p.kw + kx) * p.NC + cc) * p.Cout + ct * CT;
                            tma_load_2d(b_base + bs * b_stage, &tm_w_hi, &full_b[bs], 0, wrow);
                            tma_load_2d(b_base + bs * b_stage + b_plane, &tm_w_lo, &full_b[bs], 0, wrow);
                            if (++bs == NSTB) { bs = 0; b_phase ^= 1; }
                        }
                    }
                    if (ASETS == 2) { if (++aset == 2) { aset = 0; a_phase ^= 1; } }
                    else a_phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // per (tap, K16 step) and output row TWO MMAs: a1 x [b1 | b2s] (N = 2 CT: the hi and lo weight planes of a ring stage are
        // contiguous rows of one K-major tile) -> [main | corr], then a2s x b1 (N = CT) -> corr.  Round 1 issued three N = CT MMAs;
        // at CT = 64 each of those reads 6 KB of operands for 32 cycles of tensor work, 1.5x the shared-memory bandwidth.
        const uint32_t idesc = idesc_f16(0, 0, 128, CT), idesc2 = idesc_f16(0, 0, 128, 2 * CT);
        const bool merged = 2 * CT <= 256;
        uint32_t a_phase = 0; int aset = 0; int bs = 0; uint32_t b_phase = 0; int acc = 0; uint32_t acc_phase = 0;
        if (p.w_res && (int)blockIdx.x < nitems) mbar_wait(&full_b[0], 0);   // resident filter bank has landed
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d0 = tmem_base + (uint32_t)(acc * 4 * CT);
            for (int cc = 0; cc < p.NC; ++cc) {
                uint8_t *aset_base = a_base + aset * R * a_row;
                uint64_t *fa = full_a + aset * MAX_ROWS, *ea = empty_a + aset * MAX_ROWS;
                for (int ky = 0; ky < p.kh; ++ky) {
                    if (ky == 0) mbar_wait(&fa[0], a_phase);
                    mbar_wait(&fa[ky + 1], a_phase);
                    for (int kx = 0; kx < p.kw; ++kx) {
                        if (!p.w_res) mbar_wait(&full_b[bs], b_phase);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (elect_one()) {
                            const uint32_t sb = smem_u32(b_base + (p.w_res ? ((ky * p.kw + kx) * p.NC + cc) : bs) * b_stage);
                            const uint64_t b_hi = umma_desc_sw64(sb), b_lo = umma_desc_sw64(sb + b_plane);
                            const bool first = (cc | ky | kx) == 0;
#pragma unroll
                            for (int r = 0; r < 2; ++r) {
                                const uint32_t sa = smem_u32(aset_base + (ky + r) * a_row) + (uint32_t)(kx * PIX_B);
                                const uint64_t a_hi = umma_desc_sw64(sa), a_lo = umma_desc_sw64(sa + a_plane);
                                const uint32_t d_main = d0 + (uint32_t)(2 * r * CT), d_corr = d_main + (uint32_t)CT;
#pragma unroll
                                for (int k = 0; k < 2; ++k) {                     // 32 channels = 2 x K16
                                    const uint64_t adv = (uint64_t)((k * 32) >> 4);
                                    if (merged) {
                                        umma_f16(d_main, a_hi + adv, b_hi + adv, idesc2, (first && k == 0) ? 0u : 1u);     // [main | corr]
                                        umma_f16(d_corr, a_lo + adv, b_hi + adv, idesc, 1u);
                                    } else {
                                        umma_f16(d_corr, a_lo + adv, b_hi + adv, idesc, (first && k == 0) ? 0u : 1u);
                                        umma_f16(d_corr, a_hi + adv, b_lo + adv, idesc, 1u);
                                        umma_f16(d_main, a_hi + adv, b_hi + adv, idesc, (first && k == 0) ? 0u : 1u);
                                    }
                                }
                            }
                            if (!p.w_res) umma_commit(&empty_b[bs]);
                            if (kx == p.kw - 1) {
                                // input row buffers whose last reader (for this chunk) was this tap row: row ky (and kh when ky == kh-1)
                                umma_commit(&ea[ky]);
                                if (ky == p.kh - 1) { umma_commit(&ea[p.kh]); if (cc == p.NC - 1) umma_commit(&tfull[acc]); }
                            }
                        }
                        __syncwarp();
                        if (!p.w_res && ++bs == NSTB) { bs = 0; b_phase ^= 1; }
                    }
                }
                if (ASETS == 2) { if (++aset == 2) { aset = 0; a_phase ^= 1; } }
                else a_phase ^= 1;
            }
            if (p.acc_sets == 2) { if (++acc == 2) { acc = 0; acc_phase ^= 1; } }
            else acc_phase ^= 1;
        }
    } else {
        // ===================== epilogue (warps 2..9): thread = output column of the tile, both rows; the two warps of a TMEM
        // lane quarter take the even / odd 32-channel chunks =====================
        const int q = warp & 3, half = (warp - 2) >> 2;
        int acc = 0; uint32_t acc_phase = 0;
        bool bad = false;
        constexpr float RS = 1.f / X2_SCALE;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int ct = item % p.items_c; const int rest = item / p.items_c;
            const int ws = rest % p.items_w, hp = (rest / p.items_w) % p.items_h, n = rest / (p.items_w * p.items_h);
            mbar_wait(&tfull[acc], acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 4 * CT);
            const int cbase = ct * CT;                                     // first output channel of this tile
            const int wcol = ws * TW + q * 32 + lane;                       // conv output column
#pragma unroll 1
            for (int c0 = 32 * half; c0 < CT; c0 += 64) {
                float v0[32], v1[32];
                {
                    uint32_t a[32], t[32];
                    tmem_ld32_nowait(lane_base + (uint32_t)c0, a);
                    tmem_ld32_nowait(lane_base + (uint32_t)(CT + c0), t);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v0[j] = fmaf(__uint_as_float(t[j]), RS, __uint_as_float(a[j]));
                    tmem_ld32_nowait(lane_base + (uint32_t)(2 * CT + c0), a);
                    tmem_ld32_nowait(lane_base + (uint32_t)(3 * CT + c0), t);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v1[j] = fmaf(__uint_as_float(t[j]), RS, __uint_as_float(a[j]));
                }
                if (p.bias) {                                                 // 32 channels = 128 bytes, 16-byte aligned (Cout % 16 == 0)
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(p.bias + cbase + c0 + j));
                        v0[j] += b4.x; v0[j + 1] += b4.y; v0[j + 2] += b4.z; v0[j + 3] += b4.w;
                        v1[j] += b4.x; v1[j + 1] += b4.y; v1[j + 2] += b4.z; v1[j + 3] += b4.w;
                    }
                }
                act_apply_vec(v0, p.act); act_apply_vec(v1, p.act);
                if (p.pool) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float mx = fmaxf(v0[j], v1[j]);
                        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                        v0[j] = mx;
                    }
                    const int wp = wcol >> 1;
                    if ((lane & 1) == 0 && hp < p.out_h && wp < p.out_w) {
                        const size_t off = (size_t)((long long)n * p.sN + (long long)hp * p.sH + (long long)wp * p.sW) + cbase + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            if (p.y) {
                                *reinterpret_cast<float4 *>(p.y + off + j) = make_float4(v0[j], v0[j + 1], v0[j + 2], v0[j + 3]);
                                *reinterpret_cast<float4 *>(p.y + off + j + 4) = make_float4(v0[j + 4], v0[j + 5], v0[j + 6], v0[j + 7]);
                            }
                            if (p.y_hi) store_planes8(p.y_hi + off + j, p.y_lo + off + j, v0 + j, bad);
                        }
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int ho = 2 * hp + r;
                        if (ho < p.out_h && wcol < p.out_w) {
                            const size_t off = (size_t)((long long)n * p.sN + (long long)ho * p.sH + (long long)wcol * p.sW) + cbase + c0;
                            const float *v = r ? v1 : v0;
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                if (p.y) {
                                    *reinterpret_cast<float4 *>(p.y + off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                                    *reinterpret_cast<float4 *>(p.y + off + j + 4) = make_float4(v[j + 4], v[j + 5], v[j + 6], v[j + 7]);
                                }
                                if (p.y_hi) store_planes8(p.y_hi + off + j, p.y_lo + off + j, v + j, bad);
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (p.acc_sets == 2) { if (++acc == 2) { acc = 0; acc_phase ^= 1; } }
            else acc_phase ^= 1;
        }
        if (bad && p.flag) atomicOr(p.flag, 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// NHWC fp16 plane [N][H][W][C] (C a multiple of 32) as a 4-D tensor map (C, W, H, N); box = 32 x box_w x 1 x 1, 64B swizzle, OOB -> 0
inline bool make_map_nhwc(CUtensorMap *map, const __half *base, uint64_t N, uint64_t H, uint64_t W, uint64_t C, uint32_t box_w) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[4] = {C, W, H, N};
    cuuint64_t strides[3] = {C * sizeof(__half), W * C * sizeof(__half), H * W * C * sizeof(__half)};
    cuuint32_t box[4] = {32, box_w, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// output-channel tile, weight-ring depth and input-row buffering for a layer; returns the dynamic shared memory the kernel needs.
// Preference: two sets of input rows (the next item's rows load under the current item's MMAs: with one set the issuer stalled for a
// TMA round trip at every item boundary), then as deep a weight ring as still fits (a tap's tile feeds only 12 MMAs).
// `nc` (32-channel input chunks; 0 = unknown) lets the planner keep the whole filter bank resident when it fits next to two sets of
// input rows and the tile covers all output channels: cfg2's conv2 (3 x 3 x 32 -> 64) = 72 KB of weights + 144 KB of rows.
inline size_t conv_tc_plan(int kh, int kw, int cout, int *ct_out, int *nstb_out, int *a_row_bytes, int *a_sets_out = nullptr, int nc = 0,
                           int *w_res_out = nullptr) {
    const int ct = cout <= 128 ? cout : (cout % 128 == 0 ? 128 : (cout % 64 == 0 ? 64 : 32));
    const int plane = ((TW + kw - 1) * PIX_B + 1023) & ~1023;
    const size_t rows = (size_t)(kh + 1) * 2 * plane, stage = (size_t)2 * ct * PIX_B, fixed = 512 + 1024, cap = 227 * 1024;
    if (w_res_out) *w_res_out = 0;
    if (nc > 0 && w_res_out && ct == cout && 2 * rows + (size_t)kh * kw * nc * stage + fixed <= cap) {
        if (ct_out) *ct_out = ct;
        if (nstb_out) *nstb_out = kh * kw * nc;
        if (a_row_bytes) *a_row_bytes = plane;
        if (a_sets_out) *a_sets_out = 2;
        *w_res_out = 1;
        return 2 * rows + (size_t)kh * kw * nc * stage + fixed;
    }
    int a_sets = 2, nstb = MAX_STB;
    if (2 * rows + 4 * stage + fixed > cap) a_sets = 1;
    while (nstb > 4 && a_sets * rows + nstb * stage + fixed > cap) --nstb;
    if (ct_out) *ct_out = ct;
    if (nstb_out) *nstb_out = nstb;
    if (a_row_bytes) *a_row_bytes = plane;
    if (a_sets_out) *a_sets_out = a_sets;
    return a_sets * rows + nstb * stage + fixed;
}

}  // namespace ctc
}  // namespace kb
