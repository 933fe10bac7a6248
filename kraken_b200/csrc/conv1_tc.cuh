// conv1_tc.cuh - the first layer of a recogniser on tcgen05: conv(Cin = 1, 3x3, stride 1, "same" padding) -> [Dropout] ->
// MaxPool 2x2/2 -> (fp16 planes for the next tensor-core layer), for sm_100a.
//   kraken/lib/vgsl/layers.py:842-860 (ActConv2D), :381-388 (MaxPool)
//
// The CUDA-core version of this group (kernels.cuh k_conv1_pool33) spends 144 FFMA2 per conv pixel and is issue bound at 0.060 ms for
// cfg2 against an HBM floor of 0.014 ms.  Here the 9-tap contraction is ONE K16 step of the tensor core:
//   M = 128 pooled pixels (tile = tph pooled rows x tpw pooled columns), K = 16 (9 taps + 7 zeros), N = Cout.
//   A: im2col patches as fp16 planes, built by the threads in shared memory - one tile PER POOL POSITION (dy, dx): row r of A[dy][dx]
//      holds the nine inputs of conv pixel (2 py + dy, 2 px + dx), so that the four conv results a pooled pixel needs arrive in the four
//      accumulator column groups of ONE TMEM lane and the 2x2 max is a register max (no shuffles, no shared-memory round trip).
//      The input is split into its planes once per pixel (x = x1 + x2s 2^-11, as everywhere), not once per tap.
//   B: [Cout hi rows | Cout lo rows] x K16, packed on the host (Exec finalize), 2 KB.
//   per position two MMAs:  a1 x [b1 | b2s] (N = 2 Cout) -> [main | corr],  a2s x b1 (N = Cout) -> corr.   8 MMAs per 128 pooled pixels.
//   D: 4 positions x 2 Cout columns per accumulator set, two sets (16 Cout <= 512 columns): the MMAs of tile i run under the epilogue of
//      tile i-1 and the patch build of tile i+1.
//   epilogue: thread = pooled pixel (TMEM lane) x half of the channels: max over the four positions of main + corr 2^-11, + bias,
//      activation (monotonic: commutes with the max, as in k_conv1_pool), split into the consumer's planes, 16-byte stores.
// Operand layout: K-major WITHOUT swizzle (K = 16 halves = two 16-byte chunks per row): chunk c of row r at (r / 8) * 256 + c * 128 +
// (r % 8) * 16, i.e. core matrices of 8 rows x 16 bytes, LBO = 128 (next K chunk), SBO = 256 (next 8 rows).
#pragma once
#include "gemm_tc.cuh"
#include "lstm_tc.cuh"

namespace kb {
namespace c1tc {

using namespace kb::tc;

constexpr int C1_THREADS = 256;
constexpr int PATCH_W = 264;                              // halves per patch row (2 * 128 + 2 used at most), multiple of 8
constexpr int PATCH_ROWS = 6;                             // 2 * tph + 2 input rows, tph <= 2
constexpr int A_TILE_B = 128 * 32;                        // one plane of one position: 128 rows x 16 halves

struct Conv1TcParams {
    const float *x; const __half *wpk; const float *bias; float *y; __half *y_hi; __half *y_lo; int *flag;
    int N, H, W, Cout, Hp, Wp, act;
    int tpw, tph;                                         // pooled columns x rows of a tile (128 x 1 or 64 x 2)
    int tiles_w, tiles_h;
};

inline size_t conv1_tc_smem(int cout) {
    return (size_t)2 * PATCH_ROWS * PATCH_W * 2 + (size_t)2 * 4 * 2 * A_TILE_B + (size_t)2 * cout * 32 + 256 + 1024;      // ~74 KB: two CTAs per SM
}

template <int C>
__global__ void __launch_bounds__(C1_THREADS, 2) k_conv1_tc(Conv1TcParams p) {
    constexpr int TMC = 8 * C;                                            // TMEM columns: ONE accumulator set; two CTAs per SM overlap each other
    constexpr int CH = C / 2;                                             // channels per thread: 16 (Cout 32) or 8 (Cout 16)
    constexpr int PE = 5;                                                 // patch elements per thread: 6 x 130 or 4 x 258 <= 5 x 256
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_align1024(smem_raw);
    uint8_t *sA = smem;                                                   // [buffer][position][plane][128 rows x 32 B]
    uint8_t *sW = sA + 2 * 4 * 2 * A_TILE_B;                              // [2 Cout rows x 32 B]
    __half *sP = reinterpret_cast<__half *>(sW + 2 * C * 32);             // [plane][row][PATCH_W]
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(sP) + 2 * PATCH_ROWS * PATCH_W * 2);
    uint64_t *tfull = bars;                                               // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = p.N * p.tiles_h * p.tiles_w;
    const int prow = 2 * p.tph + 2, pcol = 2 * p.tpw + 2;

    if (tid == 0) {
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 2 * C * 32 / 16; i += C1_THREADS) reinterpret_cast<uint4 *>(sW)[i] = __ldg(reinterpret_cast<const uint4 *>(p.wpk) + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t idesc2 = idesc_f16(0, 0, 128, 2 * C), idesc1 = idesc_f16(0, 0, 128, C);

    // roles: epilogue = TMEM lane quarter q x channel half hsel; im2col = pooled pixel r x plane
    const int q = warp & 3, hsel = warp >> 2;
    const int er = 32 * q + lane, ery = er / p.tpw, erx = er - ery * p.tpw;       // epilogue pixel inside the tile
    const int ir = tid & 127, iplane = tid >> 7, iry = ir / p.tpw, irx = ir - iry * p.tpw;
    const __half *ipp = sP + (size_t)iplane * PATCH_ROWS * PATCH_W + (2 * iry) * PATCH_W + 2 * irx;
    const uint32_t ia_off = (uint32_t)(iplane * A_TILE_B + (ir >> 3) * 256 + (ir & 7) * 16);
    bool bad = false;
    constexpr float RS = 1.f / X2_SCALE;
    float bv[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) bv[j] = p.bias ? __ldg(p.bias + hsel * CH + j) : 0.f;
    // patch elements of this thread: (row j, column k) of the input window - the same for every tile
    int pj[PE], pk[PE], po[PE]; bool pv[PE];
#pragma unroll
    for (int e = 0; e < PE; ++e) {
        const int i = tid + e * C1_THREADS;
        pv[e] = i < prow * pcol; pj[e] = pv[e] ? i / pcol : 0; pk[e] = pv[e] ? i - pj[e] * pcol : 0;
        po[e] = pv[e] ? pj[e] * PATCH_W + pk[e] : (PATCH_ROWS - 1) * PATCH_W + PATCH_W - 1;      // surplus elements land in an unused column
    }

    struct Tile { int n, py0, px0; };
    auto coords = [&](int tile) {
        Tile t; const int tw = tile % p.tiles_w; const int rest = tile / p.tiles_w;
        t.py0 = (rest % p.tiles_h) * p.tph; t.n = rest / p.tiles_h; t.px0 = tw * p.tpw; return t;
    };
    // the tile after `t` in this CTA's sequence (stride gridDim.x), without divisions: the stride as (images, tile rows, tile columns)
    const Tile stride = coords((int)gridDim.x);
    auto advance = [&](const Tile &t) {
        Tile u; u.px0 = t.px0 + stride.px0; u.py0 = t.py0 + stride.py0; u.n = t.n + stride.n;
        if (u.px0 >= p.tiles_w * p.tpw) { u.px0 -= p.tiles_w * p.tpw; u.py0 += p.tph; }
        if (u.py0 >= p.tiles_h * p.tph) { u.py0 -= p.tiles_h * p.tph; u.n += 1; }
        return u;
    };
    auto patch_fetch = [&](const Tile &t, float (&v)[PE]) {                // global loads only: they fly under the epilogue
        const float *img = p.x + (size_t)t.n * p.H * p.W;
        const int y0 = 2 * t.py0 - 1, x0 = 2 * t.px0 - 1;
#pragma unroll
        for (int e = 0; e < PE; ++e) {
            const int yy = y0 + pj[e], xx = x0 + pk[e];
            v[e] = (pv[e] && (unsigned)yy < (unsigned)p.H && (unsigned)xx < (unsigned)p.W) ? __ldg(img + (size_t)yy * p.W + xx) : 0.f;   // zero = the conv's padding
        }
    };
    auto patch_store = [&](const float (&v)[PE]) {                         // -> fp16 planes in shared memory
#pragma unroll
        float amax = 0.f;
#pragma unroll
        for (int e = 0; e < PE; ++e) {
            const __half h1 = __float2half_rn(v[e]);
            const __half h2 = __float2half_rn((v[e] - __half2float(h1)) * X2_SCALE);
            amax = fmaxf(amax, fabsf(v[e]));
            sP[po[e]] = h1; sP[PATCH_ROWS * PATCH_W + po[e]] = h2;
        }
        bad |= !(amax <= 65504.f);
    };
    auto epilogue = [&](int it, const Tile &t) {
        mbar_wait(&tfull[0], (uint32_t)(it & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int py = t.py0 + ery, px = t.px0 + erx;
        const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hsel * CH);
        float vmax[CH];
#pragma unroll
        for (int pp2 = 0; pp2 < 4; pp2 += 2) {                             // two positions per TMEM round trip
            uint32_t a[2][CH], c[2][CH];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (CH == 16) {
                    ltc::tmem_ld16_nowait(tb + (uint32_t)((pp2 + u) * 2 * C), a[u]);
                    ltc::tmem_ld16_nowait(tb + (uint32_t)((pp2 + u) * 2 * C + C), c[u]);
                } else {
                    ltc::tmem_ld8_nowait(tb + (uint32_t)((pp2 + u) * 2 * C), a[u]);
                    ltc::tmem_ld8_nowait(tb + (uint32_t)((pp2 + u) * 2 * C + C), c[u]);
                }
            }
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const float v = fmaf(__uint_as_float(c[u][j]), RS, __uint_as_float(a[u][j]));
                    vmax[j] = (pp2 + u) ? fmaxf(vmax[j], v) : v;
                }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (py < p.Hp && px < p.Wp) {
#pragma unroll
            for (int j = 0; j < CH; ++j) vmax[j] += bv[j];
            act_apply_vec(vmax, p.act);
            const size_t off = (((size_t)t.n * p.Hp + py) * p.Wp + px) * C + hsel * CH;
#pragma unroll
            for (int j = 0; j < CH; j += 8) {
                if (p.y) {
                    *reinterpret_cast<float4 *>(p.y + off + j) = make_float4(vmax[j], vmax[j + 1], vmax[j + 2], vmax[j + 3]);
                    *reinterpret_cast<float4 *>(p.y + off + j + 4) = make_float4(vmax[j + 4], vmax[j + 5], vmax[j + 6], vmax[j + 7]);
                }
                if (p.y_hi) { bool ignore = false; store_planes8(p.y_hi + off + j, p.y_lo + off + j, vmax + j, ignore); }
            }
            float amax = 0.f;
#pragma unroll
            for (int j = 0; j < CH; ++j) amax = fmaxf(amax, fabsf(vmax[j]));
            bad |= !(amax <= 65504.f);
        }
    };

    int it = 0;
    Tile cur = coords(min((int)blockIdx.x, max(ntiles - 1, 0)));
    if ((int)blockIdx.x < ntiles) {
        float v0[PE];
        patch_fetch(cur, v0);
        patch_store(v0);
    }
    __syncthreads();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        // ---- 1. im2col from the patch in shared memory: this thread's 4 x 4 input neighbourhood (one plane) feeds its pixel's row in the
        // four positions' A tiles.  The halves stay packed in pairs; byte permutes pick the cross-pair combinations.
        {
            uint32_t n0[4], n1[4];                                         // row j: halves (c0, c1) and (c2, c3)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                n0[j] = *reinterpret_cast<const uint32_t *>(ipp + j * PATCH_W);
                n1[j] = *reinterpret_cast<const uint32_t *>(ipp + j * PATCH_W + 2);
            }
            uint8_t *ab = sA + (size_t)((it & 1) * 4 * 2) * A_TILE_B + ia_off;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                // per row of the 3 x 3 window: dx = 0 -> (c0, c1), c2;  dx = 1 -> (c1, c2), c3
                const uint32_t m0 = __byte_perm(n0[dy], n1[dy], 0x5432), m1 = __byte_perm(n0[dy + 1], n1[dy + 1], 0x5432), m2 = __byte_perm(n0[dy + 2], n1[dy + 2], 0x5432);
                {   // dx = 0: t = r0(c0 c1 c2) r1(c0 c1 c2) r2(c0 c1 c2)
                    uint8_t *dst = ab + (size_t)((dy * 2 + 0) * 2) * A_TILE_B;
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(n0[dy], __byte_perm(n1[dy], n0[dy + 1], 0x5410), m1, n0[dy + 2]);
                    *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(n1[dy + 2] & 0xffffu, 0u, 0u, 0u);
                }
                {   // dx = 1: t = r0(c1 c2 c3) r1(c1 c2 c3) r2(c1 c2 c3)
                    uint8_t *dst = ab + (size_t)((dy * 2 + 1) * 2) * A_TILE_B;
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(m0, __byte_perm(n1[dy], n0[dy + 1], 0x7632), n1[dy + 1], m2);
                    *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(n1[dy + 2] >> 16, 0u, 0u, 0u);
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes of A -> UMMA reads
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                                   // A complete; the patch may be overwritten
        // ---- 2. MMAs of this tile (one thread); they run under the epilogue of the previous tile
        if (warp == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t a0 = smem_u32(sA + (size_t)((it & 1) * 4 * 2) * A_TILE_B), w0 = smem_u32(sW);
                const uint64_t bd = ltc::umma_desc_nosw(w0, 128u, 256u);
#pragma unroll
                for (int pos = 0; pos < 4; ++pos) {
                    const uint32_t d = tmem_base + (uint32_t)(pos * 2 * C);
                    const uint64_t ahi = ltc::umma_desc_nosw(a0 + (uint32_t)((pos * 2) * A_TILE_B), 128u, 256u);
                    const uint64_t alo = ltc::umma_desc_nosw(a0 + (uint32_t)((pos * 2 + 1) * A_TILE_B), 128u, 256u);
                    umma_f16(d, ahi, bd, idesc2, 0u);                       // a1  x [b1 | b2s] -> [main | corr]
                    umma_f16(d + (uint32_t)C, alo, bd, idesc1, 1u);         // a2s x b1         -> corr
                }
                umma_commit(&tfull[0]);
            }
            __syncwarp();
        }
        // ---- 3. the next tile's input window: global loads in flight under the epilogue of the previous tile
        const int next = tile + (int)gridDim.x;
        float vn[PE];
        Tile nt = cur;
        if (next < ntiles) { nt = advance(cur); patch_fetch(nt, vn); }
        epilogue(it, cur);                                                 // waits for this tile's MMAs; the SM's other CTA fills the gap
        if (next < ntiles) patch_store(vn);
        __syncthreads();                                                   // patch of the next tile complete; accumulators read
        cur = nt;
    }
    if (bad && p.flag) atomicOr(p.flag, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMC) : "memory");
    }
}

// host: weights [Cout][9] (torch layout [Cout][1][3][3]) -> the B operand [b1 rows | b2s rows] x K16 in the no-swizzle layout above
inline bool pack_conv1_weights(const std::vector<float> &w, int cout, std::vector<__half> &out) {
    out.assign((size_t)2 * cout * 16, __float2half_rn(0.f));
    for (int pl = 0; pl < 2; ++pl)
        for (int co = 0; co < cout; ++co)
            for (int k = 0; k < 9; ++k) {
                const float x = w[(size_t)co * 9 + k];
                if (!(std::fabs(x) <= 65504.f)) return false;
                const __half h1 = __float2half_rn(x);
                const __half v = pl == 0 ? h1 : __float2half_rn((x - __half2float(h1)) * X2_SCALE);
                const int n = pl * cout + co;
                out[((size_t)(n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16) / 2 + (k % 8)] = v;
            }
    return true;
}

}  // namespace c1tc
}  // namespace kb
