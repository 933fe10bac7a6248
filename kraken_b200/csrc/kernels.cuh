// kernels.cuh - sm_100a device code of the kraken_b200 engine (fp32-exact path).
//
// Activations are NHWC fp32 in HBM ("pixel-major, feature-minor"): that is the layout in which
//   * a VGSL conv is an implicit GEMM  [N*Ho*Wo pixels] x [kh*kw*Cin] x [Cout]   with unit-stride K,
//   * the LSTM input projection and the output Linear are plain per-pixel GEMMs,
//   * an `Lbx`/`Lby` recurrence walks pixels with a constant stride,
//   * `S1(1x0)1,3` (fold H into the feature axis) is a permutation of whole feature rows.
// The reference is NCHW (kraken/lib/vgsl/layers.py:34-35); conversion happens once at the ABI edge.
//
// Arithmetic: everything here is IEEE fp32 FMA with accurate expf/tanhf, because CTC label sequences
// have to be bit-identical to the fp32 reference and a random-weight model has top-2 logit gaps down to
// 1e-4 (SURVEY.md 7 "hard parts").  The tensor-core GEMM (split fp16 operands on tcgen05) lives in
// gemm_tc.cuh and replaces k_conv_gemm for the shapes it supports.
#pragma once
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace kb {
namespace cg = cooperative_groups;

enum { DACT_LINEAR = 0, DACT_SIGMOID_LOGITS = 1, DACT_TANH = 2, DACT_SOFTMAX = 3, DACT_RELU = 4, DACT_LEAKY = 5 };

__device__ __forceinline__ float act_apply(float v, int act) {
    switch (act) {
    case DACT_RELU: return v > 0.f ? v : 0.f;                 // torch.relu: NaN propagates through the compare as in ATen
    case DACT_TANH: return tanhf(v);
    case DACT_LEAKY: return v > 0.f ? v : 0.01f * v;          // nn.LeakyReLU() default slope (layers.py:821)
    default: return v;                                        // linear, sigmoid-as-logits (layers.py:850-852), softmax (separate pass)
    }
}
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }
// activation over a register array with the kind test hoisted out of the element loop (one branch per call, not per element:
// the per-element switch made the tcgen05 epilogues branch- and instruction-fetch bound, ncu r01d)
template <int N>
__device__ __forceinline__ void act_apply_vec(float (&v)[N], int act) {
    if (act == DACT_RELU) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = v[i] > 0.f ? v[i] : 0.f;
    } else if (act == DACT_TANH) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = tanhf(v[i]);
    } else if (act == DACT_LEAKY) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = v[i] > 0.f ? v[i] : 0.01f * v[i];
    }
}
// Blackwell packed fp32 FMA: (acc.lo, acc.hi) += (w.lo, w.hi) * s.  ptxas folds the {s, s} pack into FFMA2's scalar-broadcast
// operand form (`FFMA2 Rd, Ra.F32x2.HI_LO, Rb.F32, Rc.F32x2.HI_LO`), so this is ONE issue slot for two IEEE fmas.
__device__ __forceinline__ void ffma2_bcast(unsigned long long &acc, unsigned long long w, float s) {
    unsigned long long sp;
    asm("mov.b64 %0, {%1, %1};" : "=l"(sp) : "f"(s));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(w), "l"(sp));
}
__device__ __forceinline__ unsigned long long pack2f(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2f(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// Tensor-core operand planes (gemm_tc.cuh): x = x1 + x2s * 2^-11 with x1 = fp16(x), x2s = fp16((x - x1) * 2^11).
// `bad` is raised when |x| leaves the fp16 range (the engine then re-runs the call on the fp32 CUDA-core kernels).
constexpr float X2_SCALE = 2048.f;
__device__ __forceinline__ void split_f16(float x, __half &h1, __half &h2, bool &bad) {
    h1 = __float2half_rn(x);
    h2 = __float2half_rn((x - __half2float(h1)) * X2_SCALE);
    bad |= fabsf(x) > 65504.f;
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) { return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16); }
// 8 consecutive channels of one pixel -> both planes (one 16-byte store each)
__device__ __forceinline__ void store_planes8(__half *hi, __half *lo, const float *v, bool &bad) {
    __half h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_f16(v[e], h[e], l[e], bad);
    *reinterpret_cast<uint4 *>(hi) = make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
    *reinterpret_cast<uint4 *>(lo) = make_uint4(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]), pack_h2(l[4], l[5]), pack_h2(l[6], l[7]));
}

// =============================================================================================
// uint8 line images -> the network's float input, on the device.  Reproduces, bit for bit,
//   v2.ToDtype(float32, scale=True)   = x.to(float32).mul_(1.0 / 255)          (torchvision to_dtype_image)
//   tensor_invert                     = im.max() - im                            (kraken/lib/functional_im_transforms.py:58-59)
//   zero right-padding to the batch width                                         (kraken/lib/vgsl/rpred.py:129-131)
// of ImageInputTransforms (kraken/lib/dataset/utils.py:148-151).  inv_max[n] >= 0: the line's maximum pixel value (the caller
// knows it from the crop; 255 for a white-padded page), < 0: no inversion.  Columns >= widths[n] become 0.
// =============================================================================================
__global__ void k_u8_lines_to_f32(const uint8_t *__restrict__ src, float *__restrict__ dst, int n_lines, int C, int H, int W,
                                  const int *__restrict__ widths, const short *__restrict__ inv_max) {
    const float s = (float)(1.0 / 255);
    const long long per = (long long)C * H * W, total = per * n_lines;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / per), w = (int)(i % W);
        float x = __fmul_rn((float)src[i], s);
        const int im = inv_max ? inv_max[n] : -1;
        if (im >= 0) x = __fsub_rn(__fmul_rn((float)im, s), x);
        if (widths && w >= widths[n]) x = 0.f;
        dst[i] = x;
    }
}

// =============================================================================================
// space-to-depth operand planes for a stride-2 convolution on the tensor cores (conv_tc.cuh): out[n][hb][wb][cs] fp16 planes with
// channel s = (h & 1, w & 1, c) for s < 4 C and zeros above / outside the image.  src is the NCHW network input (nchw = 1) or an
// NHWC activation.  One thread = 8 consecutive s2d channels of one block pixel (one 16-byte store per plane).
// =============================================================================================
__global__ void k_s2d_planes(const float *__restrict__ src, int nchw, __half *__restrict__ hi, __half *__restrict__ lo, int N, int C, int H,
                             int W, int Hb, int Wb, int cs, int *flag) {
    const int groups = cs >> 3;
    const long long total = (long long)N * Hb * Wb * groups;
    bool bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % groups); long long t = i / groups;
        const int wb = (int)(t % Wb); t /= Wb;
        const int hb = (int)(t % Hb); const int n = (int)(t / Hb);
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int sc = g * 8 + e;
            float x = 0.f;
            if (sc < 4 * C) {
                const int phs = sc / C, c = sc - phs * C, h = 2 * hb + (phs >> 1), w = 2 * wb + (phs & 1);
                if (h < H && w < W) x = nchw ? __ldg(src + (((size_t)n * C + c) * H + h) * W + w) : __ldg(src + (((size_t)n * H + h) * W + w) * C + c);
            }
            v[e] = x;
        }
        store_planes8(hi + (size_t)i * 8, lo + (size_t)i * 8, v, bad);
    }
    if (bad) atomicOr(flag, 1);
}

// =============================================================================================
// batched transpose  in[B][R][C] -> out[B][C][R]   (NCHW <-> NHWC at the ABI edge)
// =============================================================================================
__global__ void k_transpose(const float *__restrict__ in, float *__restrict__ out, int R, int C) {
    __shared__ float tile[32][33];
    const size_t b = blockIdx.z;
    const float *src = in + b * (size_t)R * C;
    float *dst = out + b * (size_t)R * C;
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        if (r < R && c < C) tile[i][threadIdx.x] = src[(size_t)r * C + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) dst[(size_t)c * R + r] = tile[threadIdx.x][i];
    }
}

// =============================================================================================
// implicit-GEMM convolution / per-pixel linear  (fp32 FFMA)
//   y[m][n] = act( sum_k A[m][k] * Wt[k][n] + bias[n] ),  m = (img, ho, wo), k = (ky, kx, ci)
// ActConv2D.forward (layers.py:842-852) incl. zero padding (dil*(k-1))//2, stride, dilation;
// LinSoftmax (layers.py:710-722) and the LSTM input projection are the kh=kw=1 case.
// =============================================================================================
struct ConvParams {
    const float *x; const float *wt; const float *bias; float *y;
    int N, H, W, Cin, Ho, Wo, Cout, Ncp;
    int kh, kw, sy, sx, dy, dx, py, px;
    int K; long long M; int act;
};

constexpr int CG_BM = 128, CG_BN = 64, CG_BK = 16, CG_NT = 256, CG_LDA = CG_BK + 4;

template <bool VEC4>
__global__ void __launch_bounds__(CG_NT, 2) k_conv_gemm(ConvParams p) {
    __shared__ __align__(16) float As[2][CG_BM][CG_LDA];
    __shared__ __align__(16) float Bs[2][CG_BK][CG_BN];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * CG_BM;
    const int n0 = blockIdx.y * CG_BN;
    const int tx = tid & 15, ty = tid >> 4;

    // ---- A-load bookkeeping: VEC4 -> 2 rows x one k-quad per thread; scalar -> 8 rows x one k per thread
    constexpr int AR = VEC4 ? 2 : 8;
    const int a_k = VEC4 ? (tid & 3) * 4 : (tid & 15);
    const int a_r0 = VEC4 ? (tid >> 2) : (tid >> 4);
    constexpr int a_rstep = VEC4 ? 64 : 16;
    long long a_base[AR]; int a_hi0[AR], a_wi0[AR]; bool a_ok[AR];
#pragma unroll
    for (int j = 0; j < AR; ++j) {
        long long m = m0 + a_r0 + j * a_rstep;
        a_ok[j] = m < p.M;
        long long mm = a_ok[j] ? m : 0;
        int wo = (int)(mm % p.Wo); long long t = mm / p.Wo;
        int ho = (int)(t % p.Ho); long long img = t / p.Ho;
        a_base[j] = img * (long long)p.H * p.W * p.Cin;
        a_hi0[j] = ho * p.sy - p.py; a_wi0[j] = wo * p.sx - p.px;
    }
    const int b_k = tid >> 4, b_n = (tid & 15) * 4;

    float4 ra4[VEC4 ? 2 : 1]; float ra1[VEC4 ? 1 : 8]; float4 rb;

    auto load_tile = [&](int kt) {
        const int k = kt * CG_BK + a_k;
        int ci = 0, kx = 0, ky = 0;
        const bool kok = k < p.K;
        if (kok) { ci = k % p.Cin; int kk = k / p.Cin; kx = kk % p.kw; ky = kk / p.kw; }
#pragma unroll
        for (int j = 0; j < AR; ++j) {
            const int hi = a_hi0[j] + ky * p.dy, wi = a_wi0[j] + kx * p.dx;
            const bool ok = kok && a_ok[j] && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            const float *src = p.x + a_base[j] + ((long long)hi * p.W + wi) * p.Cin + ci;
            if constexpr (VEC4) ra4[j] = ok ? __ldg(reinterpret_cast<const float4 *>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
            else ra1[j] = ok ? __ldg(src) : 0.f;
        }
        const int kb = kt * CG_BK + b_k;
        rb = kb < p.K ? __ldg(reinterpret_cast<const float4 *>(p.wt + (size_t)kb * p.Ncp + n0 + b_n)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int j = 0; j < AR; ++j) {
            if constexpr (VEC4) *reinterpret_cast<float4 *>(&As[buf][a_r0 + j * a_rstep][a_k]) = ra4[j];
            else As[buf][a_r0 + j * a_rstep][a_k] = ra1[j];
        }
        *reinterpret_cast<float4 *>(&Bs[buf][b_k][b_n]) = rb;
    };

    unsigned long long acc01[8], acc23[8];       // 8 rows x 4 columns as FFMA2 pairs
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc01[i] = 0ull; acc23[i] = 0ull; }

    const int nk = (p.K + CG_BK - 1) / CG_BK;
    load_tile(0); store_tile(0); __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) load_tile(kt + 1);
#pragma unroll
        for (int kk = 0; kk < CG_BK; kk += 4) {
            float4 a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4 *>(&As[cur][ty + 16 * i][kk]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(&Bs[cur][kk + j][tx * 4]);   // (b0,b1), (b2,b3)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float av = j == 0 ? a[i].x : j == 1 ? a[i].y : j == 2 ? a[i].z : a[i].w;
                    ffma2_bcast(acc01[i], b.x, av); ffma2_bcast(acc23[i], b.y, av);
                }
            }
        }
        if (kt + 1 < nk) store_tile(cur ^ 1);
        __syncthreads();
    }
    // ---- epilogue: bias + activation, NHWC store
    const int n = n0 + tx * 4;
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.bias) {
#pragma unroll
        for (int c = 0; c < 4; ++c) if (n + c < p.Cout) bv[c] = __ldg(p.bias + n + c);
    }
    const bool vec_store = (p.Cout & 3) == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + ty + 16 * i;
        if (m >= p.M) continue;
        float v[4];
        unpack2f(acc01[i], v[0], v[1]); unpack2f(acc23[i], v[2], v[3]);
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = act_apply(v[c] + bv[c], p.act);
        float *dst = p.y + (size_t)m * p.Cout + n;
        if (vec_store) { if (n < p.Cout) *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]); }
        else {
#pragma unroll
            for (int c = 0; c < 4; ++c) if (n + c < p.Cout) dst[c] = v[c];
        }
    }
}

// =============================================================================================
// fused  conv(Cin = 1, stride 1) + bias + activation + 2x2/2 max-pool   (ActConv2D -> [Dropout] -> MaxPool)
// The first layer of every kraken recogniser reads a 1-channel line image: it is a stencil, not a GEMM, and its
// full-resolution output (cfg2: 315 MB per batch) is only ever consumed by the pool that follows.  One block = 64 pooled
// pixels of one pooled row x all output channels; the (kh+1) x (128+kw-1) input patch and the filter bank sit in shared
// memory; each thread produces 8 channels of one pooled pixel (4 conv positions).  Optionally also writes the fp16 operand
// planes of the result for a tensor-core consumer.  max(relu(a), relu(b)) == relu(max(a, b)): pooling first is exact.
// =============================================================================================
struct Conv1PoolParams {
    const float *x; const float *wt; const float *bias; float *y; __half *y_hi; __half *y_lo; int *flag;
    int N, H, W, Cout, Ncp, kh, kw, py, px, Hp, Wp, act;
};
__global__ void __launch_bounds__(256) k_conv1_pool(Conv1PoolParams p) {
    extern __shared__ float c1_sm[];
    const int cgroups = p.Cout >> 3;                                   // 8 channels per thread
    const int ppb = 256 / cgroups;                                     // pooled pixels per block
    const int tw = 2 * ppb + p.kw - 1, th = p.kh + 1;                  // input patch
    float *s_in = c1_sm;                                               // [th][tw]
    float *s_w = c1_sm + ((th * tw + 3) & ~3);                         // [kh*kw][Cout], 16-byte aligned
    const int n = blockIdx.z, hp = blockIdx.y, wp0 = blockIdx.x * ppb;
    const int tid = threadIdx.x;
    for (int i = tid; i < p.kh * p.kw * p.Cout; i += 256) s_w[i] = __ldg(p.wt + (size_t)(i / p.Cout) * p.Ncp + (i % p.Cout));
    const int h0 = 2 * hp - p.py, w0 = 2 * wp0 - p.px;
    const float *img = p.x + (size_t)n * p.H * p.W;
    for (int i = tid; i < th * tw; i += 256) {
        const int r = i / tw, c = i % tw, hi = h0 + r, wi = w0 + c;
        s_in[i] = (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) ? __ldg(img + (size_t)hi * p.W + wi) : 0.f;
    }
    __syncthreads();
    const int cg_i = tid % cgroups, px = tid / cgroups, wp = wp0 + px;
    if (wp >= p.Wp) return;
    unsigned long long acc2[4][4];                 // 4 conv positions x 8 channels as FFMA2 pairs
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc2[q][c] = 0ull;
    for (int ky = 0; ky < p.kh; ++ky)
        for (int kx = 0; kx < p.kw; ++kx) {
            const float *wv = s_w + (ky * p.kw + kx) * p.Cout + cg_i * 8;
            const ulonglong2 wa = *reinterpret_cast<const ulonglong2 *>(wv), wb = *reinterpret_cast<const ulonglong2 *>(wv + 4);
            const float *ip = s_in + ky * tw + 2 * px + kx;
            const float in4[4] = {ip[0], ip[1], ip[tw], ip[tw + 1]};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ffma2_bcast(acc2[q][0], wa.x, in4[q]); ffma2_bcast(acc2[q][1], wa.y, in4[q]);
                ffma2_bcast(acc2[q][2], wb.x, in4[q]); ffma2_bcast(acc2[q][3], wb.y, in4[q]);
            }
        }
    float acc[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < 4; ++c) unpack2f(acc2[q][c], acc[q][2 * c], acc[q][2 * c + 1]);
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float b = p.bias ? __ldg(p.bias + cg_i * 8 + c) : 0.f;
        const float m = fmaxf(fmaxf(acc[0][c], acc[1][c]), fmaxf(acc[2][c], acc[3][c])) + b;   // + b commutes with max
        o[c] = act_apply(m, p.act);
    }
    const size_t off = (((size_t)n * p.Hp + hp) * p.Wp + wp) * p.Cout + cg_i * 8;
    if (p.y) {            // the fp32 tensor is skipped when the only consumer reads the fp16 planes
        *reinterpret_cast<float4 *>(p.y + off) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4 *>(p.y + off + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
    if (p.y_hi) {
        bool bad = false;
        store_planes8(p.y_hi + off, p.y_lo + off, o, bad);
        if (bad) atomicOr(p.flag, 1);
    }
}

// 3x3 specialisation of the fused stencil (cfg2's first layer): the filter bank lives in REGISTERS (9 taps x 8 channels = 36
// packed pairs per thread) and a block walks a strip of RP pooled rows, so per pooled pixel a thread issues 8 LDS.64 for its
// 4x4 input patch, 144 FFMA2 and the epilogue; the generic kernel above reloads the bank per block and reads 2 LDS.128 of
// weights per tap (measured 0.107 ms on cfg2 = 5x its FMA bound).
template <int RP, int NE>                     // NE = ceil(4 * tw / 256): patch elements per thread
__global__ void __launch_bounds__(256) k_conv1_pool33(Conv1PoolParams p) {
    extern __shared__ float c1_sm[];                                   // [2][4][tw]
    const int cgroups = p.Cout >> 3, ppb = 256 / cgroups;
    const int tw = 2 * ppb + 2;
    const int n = blockIdx.z, wp0 = blockIdx.x * ppb;
    const int tid = threadIdx.x, cg_i = tid % cgroups, px = tid / cgroups, wp = wp0 + px;
    unsigned long long wr[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p.wt + (size_t)t * p.Ncp + cg_i * 8));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(p.wt + (size_t)t * p.Ncp + cg_i * 8 + 4));
        wr[t][0] = pack2f(a.x, a.y); wr[t][1] = pack2f(a.z, a.w); wr[t][2] = pack2f(b.x, b.y); wr[t][3] = pack2f(b.z, b.w);
    }
    float bias[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bias[c] = p.bias ? __ldg(p.bias + cg_i * 8 + c) : 0.f;
    const float *img = p.x + (size_t)n * p.H * p.W;
    const int w0 = 2 * wp0 - p.px;
    const int hp_end = min(p.Hp, ((int)blockIdx.y + 1) * RP);
    bool bad = false;
    int buf = 0;
    // the 4 x tw input patch of a pooled row: NE elements per thread, fetched one row AHEAD into registers so that the global
    // latency hides under the previous row's FMAs
    float pre[NE];
    auto fetch = [&](int hp) {
        const int h0 = 2 * hp - p.py;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const int i = tid + e * 256;
            const int r = (i >= tw) + (i >= 2 * tw) + (i >= 3 * tw), c = i - r * tw, hi = h0 + r, wi = w0 + c;
            pre[e] = (i < 4 * tw && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) ? __ldg(img + (size_t)hi * p.W + wi) : 0.f;
        }
    };
    fetch(blockIdx.y * RP);
    for (int hp = blockIdx.y * RP; hp < hp_end; ++hp, buf ^= 1) {
        float *s_in = c1_sm + buf * 4 * tw;
#pragma unroll
        for (int e = 0; e < NE; ++e) { const int i = tid + e * 256; if (i < 4 * tw) s_in[i] = pre[e]; }
        __syncthreads();
        if (hp + 1 < hp_end) fetch(hp + 1);
        if (wp < p.Wp) {
            float v[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float2 a = *reinterpret_cast<const float2 *>(s_in + r * tw + 2 * px), b = *reinterpret_cast<const float2 *>(s_in + r * tw + 2 * px + 2);
                v[r][0] = a.x; v[r][1] = a.y; v[r][2] = b.x; v[r][3] = b.y;
            }
            unsigned long long acc2[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc2[q][c] = 0ull;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float in4[4] = {v[ky][kx], v[ky][kx + 1], v[ky + 1][kx], v[ky + 1][kx + 1]};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int c = 0; c < 4; ++c) ffma2_bcast(acc2[q][c], wr[ky * 3 + kx][c], in4[q]);
                }
            float o[8];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float a0, a1, b0, b1, c0, c1, d0, d1;
                unpack2f(acc2[0][c], a0, a1); unpack2f(acc2[1][c], b0, b1); unpack2f(acc2[2][c], c0, c1); unpack2f(acc2[3][c], d0, d1);
                o[2 * c] = fmaxf(fmaxf(a0, b0), fmaxf(c0, d0)) + bias[2 * c];            // + b commutes with max
                o[2 * c + 1] = fmaxf(fmaxf(a1, b1), fmaxf(c1, d1)) + bias[2 * c + 1];
            }
            act_apply_vec(o, p.act);
            const size_t off = (((size_t)n * p.Hp + hp) * p.Wp + wp) * p.Cout + cg_i * 8;
            if (p.y) {
                *reinterpret_cast<float4 *>(p.y + off) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4 *>(p.y + off + 4) = make_float4(o[4], o[5], o[6], o[7]);
            }
            if (p.y_hi) store_planes8(p.y_hi + off, p.y_lo + off, o, bad);
        }
    }
    if (bad) atomicOr(p.flag, 1);
}

// softmax over the feature axis of NHWC rows, in place ('m' convs, layers.py:817-818). One warp per pixel.
__global__ void k_softmax_rows(float *__restrict__ x, long long rows, int C) {
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float *r = x + (size_t)row * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, r[c]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(r[c] - m);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    for (int c = lane; c < C; c += 32) r[c] = expf(r[c] - m) / s;
}

// =============================================================================================
// max pooling, NHWC, floor mode, no padding (layers.py:379-388)
// =============================================================================================
__global__ void k_maxpool(const float *__restrict__ x, float *__restrict__ y, int N, int H, int W, int C,
                          int Ho, int Wo, int kh, int kw, int sy, int sx) {
    const int cv = (C & 3) == 0 ? 4 : 1;
    const int Cq = C / cv;
    const long long total = (long long)N * Ho * Wo * Cq;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int cq = (int)(idx % Cq); long long t = idx / Cq;
        int wo = (int)(t % Wo); t /= Wo;
        int ho = (int)(t % Ho); long long n = t / Ho;
        const float *base = x + ((n * H + (long long)ho * sy) * W + (long long)wo * sx) * C + cq * cv;
        if (cv == 4) {
            float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            for (int i = 0; i < kh; ++i)
                for (int j = 0; j < kw; ++j) {
                    float4 v = __ldg(reinterpret_cast<const float4 *>(base + ((long long)i * W + j) * C));
                    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
                }
            *reinterpret_cast<float4 *>(y + idx * 4) = m;
        } else {
            float m = -INFINITY;
            for (int i = 0; i < kh; ++i)
                for (int j = 0; j < kw; ++j) m = fmaxf(m, __ldg(base + ((long long)i * W + j) * C));
            y[idx] = m;
        }
    }
}

// =============================================================================================
// generic Reshape (layers.py:313-333) as a gather on NHWC storage with NCHW index semantics
// =============================================================================================
struct ReshapeParams {
    long long in_dims[4], out_dims[4];  // NCHW
    long long shape5[5];                // split input shape
    int perm[5];                        // permuted[i] = shape5[perm[i]]
    int src, dest;
};
__global__ void k_reshape(const float *__restrict__ x, float *__restrict__ y, ReshapeParams p) {
    const long long oN = p.out_dims[0], oC = p.out_dims[1], oH = p.out_dims[2], oW = p.out_dims[3];
    const long long total = oN * oC * oH * oW;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        // idx enumerates the NHWC output: c fastest
        long long o[4];
        o[1] = idx % oC; long long t = idx / oC;
        o[3] = t % oW; t /= oW;
        o[2] = t % oH; o[0] = t / oH;
        // 4-D out coord -> 5-D permuted coord
        long long q[5]; int j = 0;
        for (int i = 0; i < 4; ++i) {
            if (i == p.dest) { long long inner = p.shape5[p.perm[i + 1]]; q[j++] = o[i] / inner; q[j++] = o[i] % inner; }
            else q[j++] = o[i];
        }
        long long s5[5];
        for (int i = 0; i < 5; ++i) s5[p.perm[i]] = q[i];
        long long in4[4]; j = 0;
        for (int i = 0; i < 4; ++i) {
            if (i == p.src) { in4[i] = s5[j] * p.shape5[j + 1] + s5[j + 1]; j += 2; }
            else in4[i] = s5[j++];
        }
        const long long src = ((in4[0] * p.in_dims[2] + in4[2]) * p.in_dims[3] + in4[3]) * p.in_dims[1] + in4[1];
        y[idx] = __ldg(x + src);
    }
}

// Addition (layers.py:205-212): out[.., j, ..] = sum_win in[.., win*chunk + j, ..] along NCHW dim `dim`
__global__ void k_addition(const float *__restrict__ x, float *__restrict__ y, long long iN, long long iC, long long iH, long long iW,
                           int dim, int chunk) {
    long long od[4] = {iN, iC, iH, iW};
    const long long D = od[dim];
    od[dim] = chunk;
    const long long nwin = D / chunk;
    const long long total = od[0] * od[1] * od[2] * od[3];
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        long long o[4];
        o[1] = idx % od[1]; long long t = idx / od[1];
        o[3] = t % od[3]; t /= od[3];
        o[2] = t % od[2]; o[0] = t / od[2];
        float s = 0.f;
        for (long long wdx = 0; wdx < nwin; ++wdx) {
            long long i4[4] = {o[0], o[1], o[2], o[3]};
            i4[dim] = wdx * chunk + o[dim];
            s += __ldg(x + ((i4[0] * iH + i4[2]) * iW + i4[3]) * iC + i4[1]);
        }
        y[idx] = s;
    }
}

// channel concat for MultiParamParallel (layers.py:60-71): y[pix][coff + c] = x[pix][c]
__global__ void k_concat(const float *__restrict__ x, float *__restrict__ y, long long pixels, int C, int Ctot, int coff) {
    const long long total = pixels * C;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        long long pix = idx / C; int c = (int)(idx % C);
        y[pix * Ctot + coff + c] = __ldg(x + idx);
    }
}

// summarising RNN (layers.py:539-541): keep the last step along W (x) or H (y)
__global__ void k_take_last(const float *__restrict__ x, float *__restrict__ y, int N, int H, int W, int C, int along_h) {
    const int oH = along_h ? 1 : H, oW = along_h ? W : 1;
    const long long total = (long long)N * oH * oW * C;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % C); long long t = idx / C;
        int w = (int)(t % oW); t /= oW;
        int h = (int)(t % oH); long long n = t / oH;
        int sh = along_h ? H - 1 : h, sw = along_h ? w : W - 1;
        y[idx] = __ldg(x + ((n * H + sh) * W + sw) * C + c);
    }
}

// =============================================================================================
// GroupNorm (layers.py:967-984): per-(sample, group) statistics over (C/G, H, W[:len]); outputs beyond
// a line's valid width are zero.  Three deterministic passes: block partials -> finalize -> apply.
// =============================================================================================
__global__ void k_gn_partial(const float *__restrict__ x, double *__restrict__ partial, int H, int W, int C, int G,
                             const int *__restrict__ lens, int chunks, int rows, int cthreads, int cpt) {
    // block = rows x cthreads threads; thread (r, ct) accumulates channels ct + j*cthreads (j < cpt) over pixels r, r+rows, ...
    extern __shared__ double gn_sm[];                        // [blockDim.x * cpt][2]
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int len = lens ? min(max(lens[n], 1), W) : W;  // seq_len.clamp(min=1, max=W)
    const long long npix = (long long)H * W;
    const long long per = (npix + chunks - 1) / chunks;
    const long long p0 = chunk * per, p1 = min(npix, p0 + per);
    const int r = threadIdx.x / cthreads, ct = threadIdx.x % cthreads;
    const float *base = x + (size_t)n * npix * C;
    for (int j = 0; j < cpt; ++j) {
        const int c = ct + j * cthreads;
        double s = 0.0, ss = 0.0;
        if (c < C && r < rows) {
            for (long long pix = p0 + r; pix < p1; pix += rows) {
                if ((int)(pix % W) < len) { const double v = (double)__ldg(base + pix * C + c); s += v; ss += v * v; }
            }
        }
        gn_sm[(threadIdx.x * cpt + j) * 2] = s; gn_sm[(threadIdx.x * cpt + j) * 2 + 1] = ss;
    }
    __syncthreads();
    const int cg_sz = C / G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double s = 0.0, ss = 0.0;
        for (int c = g * cg_sz; c < (g + 1) * cg_sz; ++c) {
            const int ctc = c % cthreads, j = c / cthreads;
            for (int rr = 0; rr < rows; ++rr) {
                const int t = rr * cthreads + ctc;
                s += gn_sm[(t * cpt + j) * 2]; ss += gn_sm[(t * cpt + j) * 2 + 1];
            }
        }
        partial[(((size_t)n * chunks + chunk) * G + g) * 2] = s;
        partial[(((size_t)n * chunks + chunk) * G + g) * 2 + 1] = ss;
    }
}
__global__ void k_gn_finalize(const double *__restrict__ partial, float2 *__restrict__ stats, int N, int G, int chunks,
                              int H, int W, int C, const int *__restrict__ lens, float eps) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * G) return;
    const int n = idx / G, g = idx % G;
    double s = 0.0, ss = 0.0;
    for (int c = 0; c < chunks; ++c) { s += partial[(((size_t)n * chunks + c) * G + g) * 2]; ss += partial[(((size_t)n * chunks + c) * G + g) * 2 + 1]; }
    const int len = lens ? min(max(lens[n], 1), W) : W;
    const double cnt = (double)(C / G) * H * len;
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[idx] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
}
__global__ void k_gn_apply(const float *__restrict__ x, float *__restrict__ y, const float2 *__restrict__ stats,
                           const float *__restrict__ gamma, const float *__restrict__ beta, long long total,
                           int H, int W, int C, int G, const int *__restrict__ lens) {
    const int cg_sz = C / G;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C); long long t = idx / C;
        const int w = (int)(t % W); const long long n = t / W / H;
        const int len = lens ? min(max(lens[n], 1), W) : W;
        float v = 0.f;
        if (w < len) {
            const float2 st = stats[n * G + c / cg_sz];
            v = (__ldg(x + idx) - st.x) * st.y * __ldg(gamma + c) + __ldg(beta + c);
        }
        y[idx] = v;
    }
}

// ---- vectorised GroupNorm (C % 4 == 0, C <= 1024): the kernels above spend their time in scalar loads and 64-bit div/mod
//      (2.97 ms for 1.24 GB on cfg3's Gn_1 = 1.25 TB/s); these stream float4 quads with the channel quad fixed per thread.
// block = 256 threads: thread -> (pixel row r = tid / C4, channel quad q = tid % C4), rows = 256 / C4 pixel rows per sweep
__global__ void __launch_bounds__(256) k_gn_stats4(const float *__restrict__ x, double *__restrict__ partial, int H, int W, int C, int G,
                                                   const int *__restrict__ lens, int chunks) {
    __shared__ double gsm[256 * 8];                          // [thread][channel of the quad][sum, sumsq]
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int C4 = C >> 2, rows = 256 / C4;
    const int r = threadIdx.x / C4, q = threadIdx.x - r * C4;
    const long long npix = (long long)H * W;
    const long long per = (npix + chunks - 1) / chunks;
    const long long p0 = chunk * per, p1 = min(npix, p0 + per);
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    if (r < rows) {
        const float4 *px = reinterpret_cast<const float4 *>(x + (size_t)n * npix * C) + q;
        if (!lens) {
            for (long long pix = p0 + r; pix < p1; pix += rows) {
                const float4 v = __ldg(px + pix * C4);
                s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
                t0 += (double)v.x * v.x; t1 += (double)v.y * v.y; t2 += (double)v.z * v.z; t3 += (double)v.w * v.w;
            }
        } else {
            const int len = min(max(lens[n], 1), W);         // seq_len.clamp(min=1, max=W)
            int w = (int)((p0 + r) % W);
            for (long long pix = p0 + r; pix < p1; pix += rows) {
                if (w < len) {
                    const float4 v = __ldg(px + pix * C4);
                    s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
                    t0 += (double)v.x * v.x; t1 += (double)v.y * v.y; t2 += (double)v.z * v.z; t3 += (double)v.w * v.w;
                }
                w += rows; while (w >= W) w -= W;
            }
        }
    }
    double *mine = gsm + threadIdx.x * 8;
    mine[0] = s0; mine[1] = t0; mine[2] = s1; mine[3] = t1; mine[4] = s2; mine[5] = t2; mine[6] = s3; mine[7] = t3;
    __syncthreads();
    const int cg_sz = C / G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double s = 0.0, ss = 0.0;
        for (int c = g * cg_sz; c < (g + 1) * cg_sz; ++c)
            for (int rr = 0; rr < rows; ++rr) {
                const double *e = gsm + ((rr * C4 + (c >> 2)) * 4 + (c & 3)) * 2;
                s += e[0]; ss += e[1];
            }
        partial[(((size_t)n * chunks + chunk) * G + g) * 2] = s;
        partial[(((size_t)n * chunks + chunk) * G + g) * 2 + 1] = ss;
    }
}
// per (sample, channel): y = x * a + b with a = rstd * gamma, b = beta - mean * a
__global__ void k_gn_coeffs(const float2 *__restrict__ stats, const float *__restrict__ gamma, const float *__restrict__ beta,
                            float2 *__restrict__ ab, int N, int C, int G) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * C) return;
    const int n = idx / C, c = idx - n * C;
    const float2 st = stats[n * G + c / (C / G)];
    const double a = (double)st.y * (double)gamma[c];
    ab[idx] = make_float2((float)a, (float)((double)beta[c] - (double)st.x * a));
}
// y (optional), y_hi / y_lo (optional fp16 planes for a tensor-core consumer)
__global__ void __launch_bounds__(256) k_gn_apply4(const float *__restrict__ x, float *__restrict__ y, __half *__restrict__ y_hi,
                                                   __half *__restrict__ y_lo, const float2 *__restrict__ ab, int H, int W, int C,
                                                   const int *__restrict__ lens, int *flag, int s2d) {
    // s2d != 0: the planes are written in space-to-depth order [n][h/2][w/2][(h&1, w&1, c)] for a stride-2 tensor-core convolution
    // (odd H / W: the missing phase of the last block row / column is written as zeros here)
    const int n = blockIdx.y;
    const int C4 = C >> 2, rows = 256 / C4;
    const int r = threadIdx.x / C4, q = threadIdx.x - r * C4;
    if (r >= rows) return;
    const long long npix = (long long)H * W;
    const float2 *pab = ab + (size_t)n * C + 4 * q;
    const float2 k0 = pab[0], k1 = pab[1], k2 = pab[2], k3 = pab[3];
    const size_t base = (size_t)n * npix * C4 + q;
    const float4 *px = reinterpret_cast<const float4 *>(x) + base;
    float4 *py = y ? reinterpret_cast<float4 *>(y) + base : nullptr;
    uint2 *ph = y_hi ? reinterpret_cast<uint2 *>(y_hi) + (s2d ? 0 : base) : nullptr;
    uint2 *pl = y_lo ? reinterpret_cast<uint2 *>(y_lo) + (s2d ? 0 : base) : nullptr;
    const int Hb = (H + 1) >> 1, Wb = (W + 1) >> 1;
    bool bad = false;
    const int len = lens ? min(max(lens[n], 1), W) : W;
    const long long stride = (long long)gridDim.x * rows;
    long long pix = (long long)blockIdx.x * rows + r;
    int w = lens ? (int)(pix % W) : 0;
    const int wstep = lens ? (int)(stride % W) : 0;
    for (; pix < npix; pix += stride) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (w < len) {
            const float4 v = __ldcs(px + pix * C4);
            o.x = fmaf(v.x, k0.x, k0.y); o.y = fmaf(v.y, k1.x, k1.y); o.z = fmaf(v.z, k2.x, k2.y); o.w = fmaf(v.w, k3.x, k3.y);
        }
        if (py) py[pix * C4] = o;
        if (ph) {
            __half h[4], l[4];
            split_f16(o.x, h[0], l[0], bad); split_f16(o.y, h[1], l[1], bad); split_f16(o.z, h[2], l[2], bad); split_f16(o.w, h[3], l[3], bad);
            const uint2 vh = make_uint2(pack_h2(h[0], h[1]), pack_h2(h[2], h[3])), vl = make_uint2(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]));
            if (!s2d) { ph[pix * C4] = vh; pl[pix * C4] = vl; }
            else {
                const int hh = (int)(pix / W), ww = (int)(pix - (long long)hh * W);
                const size_t blk = (((size_t)n * Hb + (hh >> 1)) * Wb + (ww >> 1)) * 4;          // 4 phases of C channels each
                const size_t d = (blk + (hh & 1) * 2 + (ww & 1)) * C4 + q;
                ph[d] = vh; pl[d] = vl;
                const uint2 z = make_uint2(0u, 0u);
                const bool last_w = (W & 1) && ww == W - 1, last_h = (H & 1) && hh == H - 1;
                if (last_w) { ph[d + C4] = z; pl[d + C4] = z; }
                if (last_h) { ph[d + 2 * C4] = z; pl[d + 2 * C4] = z; }
                if (last_w && last_h) { ph[d + 3 * C4] = z; pl[d + 3 * C4] = z; }
            }
        }
        if (lens) { w += wstep; if (w >= W) w -= W; }
    }
    if (bad) atomicOr(flag, 1);
}

// =============================================================================================
// LSTM recurrence (TransposedSummarizingRNN / nn.LSTM, layers.py:513-547), one direction per blockIdx.y.
//
// Input  gx  : per-pixel gate pre-activations x_t W_ih^T + b_ih + b_hh from the projection GEMM,
//              feature order [dir][unit][gate i,f,g,o]  (one float4 per (unit, pixel)).
// Output out : per-pixel hidden states, feature order [dir][unit] (forward | reverse halves).
// A thread-block cluster of CS = KSPLIT CTAs owns BL = 64/KSPLIT sequences of one direction for all
// time steps.  W_hh lives in REGISTERS for the whole kernel: thread (unit, k-slice) holds the 4 gate rows
// x 32 k of its unit (128 registers); h_{t-1} is staged in shared memory in every CTA of the cluster and
// exchanged through distributed shared memory once per step.  fp32 cell state, accurate expf/tanhf.
// Packed-sequence semantics (pack_padded_sequence, layers.py:528-536): line q runs len[q] steps, the
// reverse direction starts at its own last valid column, outputs beyond len are zero.
// =============================================================================================
// Generic recurrence for hidden sizes no resident-weight kernel takes (> 256): one step = one fp32 GEMM launch
// (G[seq][4h] = H[seq][h] . W_hh^T, k_conv_gemm) + this pointwise kernel.  2 T launches per layer, any size, same arithmetic as
// k_lstm_rec (expf / tanhf, fp32 cell state, packed-sequence semantics: a line stops at its own length, the reverse direction starts
// at each line's own end, padded steps are zero).  Layout as everywhere: gate columns [unit][i, f, g, o] (per direction).
struct LstmStepParams {
    const float *G; const float *gx; float *hstate; float *cstate; float *out; const int *lens;
    int nseq, T, hid, dirs, dir, t;                // t = step counter (0 .. maxlen-1)
    int q2; long long s_outer, s_inner, step;
    const float *peep;                             // ocropy cell (layers.py:74-103): [ip | fp | op][hid] of this direction, else NULL
};
__global__ void k_lstm_generic_step(LstmStepParams p) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.nseq * p.hid) return;
    const int u = (int)(idx % p.hid), q = (int)(idx / p.hid);
    const int len = p.lens ? min(max(p.lens[q], 0), p.T) : p.T;
    if (p.t >= len) return;                        // finished line: its state is never read again, its padded outputs were zeroed up front
    const int tt = p.dir ? len - 1 - p.t : p.t;
    const long long pix = (long long)(q / p.q2) * p.s_outer + (long long)(q % p.q2) * p.s_inner + (long long)tt * p.step;
    const float4 g4 = *reinterpret_cast<const float4 *>(p.G + ((size_t)q * p.hid + u) * 4);
    const float4 x4 = *reinterpret_cast<const float4 *>(p.gx + (size_t)pix * (p.dirs * 4 * p.hid) + (size_t)p.dir * 4 * p.hid + (size_t)u * 4);
    float c, h;
    if (!p.peep) {
        const float gi = 1.f / (1.f + expf(-(g4.x + x4.x))), gf = 1.f / (1.f + expf(-(g4.y + x4.y)));
        const float gg = tanhf(g4.z + x4.z), go = 1.f / (1.f + expf(-(g4.w + x4.w)));
        c = gf * p.cstate[idx] + gi * gg;
        h = go * tanhf(c);
    } else {
        // PeepholeLSTMCell: peepholes from c_{t-1} into input / forget gate, from c_t into the output gate - which the reference does NOT
        // squash (hy = (outgate + w_op * cy) * tanh(cy), layers.py:98-101)
        const float cx = p.cstate[idx];
        const float gi = 1.f / (1.f + expf(-((g4.x + x4.x) + p.peep[u] * cx))), gf = 1.f / (1.f + expf(-((g4.y + x4.y) + p.peep[p.hid + u] * cx)));
        const float gg = tanhf(g4.z + x4.z);
        c = gf * cx + gi * gg;
        h = ((g4.w + x4.w) + p.peep[2 * p.hid + u] * c) * tanhf(c);
    }
    p.cstate[idx] = c; p.hstate[idx] = h;
    p.out[(size_t)pix * (p.dirs * p.hid) + (size_t)p.dir * p.hid + u] = h;
}
// zero the output pixels beyond each line's length (both directions) and the initial h / c state
__global__ void k_lstm_generic_init(float *out, float *hstate, float *cstate, const int *lens, int nseq, int T, int hid, int dirs, int q2,
                                    long long s_outer, long long s_inner, long long step) {
    const long long total = (long long)nseq * T * dirs * hid;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % (dirs * hid)); const long long r = i / (dirs * hid);
        const int t = (int)(r % T), q = (int)(r / T);
        const int len = lens ? min(max(lens[q], 0), T) : T;
        if (t >= len) {
            const long long pix = (long long)(q / q2) * s_outer + (long long)(q % q2) * s_inner + (long long)t * step;
            out[(size_t)pix * (dirs * hid) + c] = 0.f;
        }
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)dirs * nseq * hid; i += (long long)gridDim.x * blockDim.x) {
        hstate[i] = 0.f; cstate[i] = 0.f;
    }
}

struct LstmParams {
    const float *gx; const float *whh; float *out; const int *lens;
    int nseq, T, hid, dirs, U;
    int q2; long long s_outer, s_inner, step;      // pixel(q, t) = (q / q2) * s_outer + (q % q2) * s_inner + t * step
};

__device__ __forceinline__ uint32_t cvta_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// remote (or local) shared-memory store that signals `bytes` on the destination CTA's mbarrier when it lands
__device__ __forceinline__ void st_async_f32(uint32_t raddr, float v, uint32_t rmbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(__float_as_uint(v)), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "LW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LD_%=;\n\t"
        "bra LW_%=;\n\t"
        "LD_%=:\n\t}" ::"r"(mbar), "r"(parity) : "memory");
}

// KSPLIT = cluster size = k-slices; BLT = lines per line-slice (8, or 10 for KSPLIT == 8 so that 64 lines x 2 directions
// need 14 clusters instead of 16 - only 15 clusters of 8 CTAs are co-resident on a B200).
template <int KSPLIT, int BLT>
__global__ void __launch_bounds__(256, 1) k_lstm_rec(LstmParams p) {
    static_assert(BLT == 8 || (BLT == 10 && KSPLIT == 8), "unsupported line blocking");
    constexpr int CS = KSPLIT, LSPLIT = 8 / KSPLIT, BL = BLT * LSPLIT, HLD = 36 * KSPLIT;
    constexpr int LPT = 8 / KSPLIT;                    // cells of lines 0..7 of a slice owned per lane after the reduce-scatter
    constexpr int NOWN = BLT == 10 ? 2 : LPT;          // + lines 8,9 owned by lanes ks = 0,1
    struct Smem { float h[2][BL][HLD]; unsigned long long mbar[2]; };
    __shared__ __align__(16) Smem sm;
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    const int ks = tid % KSPLIT, ls = (tid / KSPLIT) % LSPLIT, rg = tid >> 3;
    const int rank = CS > 1 ? (int)cluster.block_rank() : 0;
    const int chunk = blockIdx.x / CS, dir = blockIdx.y;
    const int u = rank * p.U + rg;
    const bool uvalid = rg < p.U && u < p.hid;
    const int hid = p.hid, GC = p.dirs * 4 * hid, OC = p.dirs * hid;

    unsigned long long w2[2][32];               // gate pairs (i,f) and (g,o) of this thread's unit, packed for FFMA2
#pragma unroll
    for (int gp = 0; gp < 2; ++gp)
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) {
            const int k = ks * 32 + kk;
            const bool ok = uvalid && k < hid;
            const float a = ok ? __ldg(p.whh + ((size_t)dir * 4 * hid + (size_t)(2 * gp) * hid + u) * hid + k) : 0.f;
            const float b = ok ? __ldg(p.whh + ((size_t)dir * 4 * hid + (size_t)(2 * gp + 1) * hid + u) * hid + k) : 0.f;
            w2[gp][kk] = pack2f(a, b);
        }
    for (int i = tid; i < 2 * BL * HLD; i += 256) (&sm.h[0][0][0])[i] = 0.f;

    // cells (unit u, line) this lane owns: slot j < LPT -> line ks*LPT + j of the slice; BLT == 10: slot 1 -> line 8 + ks (ks < 2)
    int len[NOWN], lline[NOWN]; long long base[NOWN]; float cst[NOWN], hlast[NOWN]; bool qvalid[NOWN];
#pragma unroll
    for (int j = 0; j < NOWN; ++j) {
        const bool slot_ok = BLT == 10 ? (j == 0 || ks < 2) : true;
        lline[j] = BLT == 10 ? (j == 0 ? ks : 8 + ks) : ls * 8 + ks * LPT + j;
        const int q = chunk * BL + lline[j];
        qvalid[j] = slot_ok && q < p.nseq;
        int l = qvalid[j] ? (p.lens ? p.lens[q] : p.T) : 0;
        len[j] = min(max(l, 0), p.T);
        const int qq = qvalid[j] ? q : 0;
        base[j] = (long long)(qq / p.q2) * p.s_outer + (long long)(qq % p.q2) * p.s_inner;
        cst[j] = 0.f; hlast[j] = 0.f;
    }
    int maxlen = 0, nvalid = 0;
    for (int lb = 0; lb < BL; ++lb) {
        const int q = chunk * BL + lb;
        if (q < p.nseq) {
            ++nvalid;
            const int l = p.lens ? min(max(p.lens[q], 0), p.T) : p.T;
            maxlen = max(maxlen, l);
            if (l < p.T) {              // zero the padded tail of this CTA's units (pad_packed_sequence pads with 0)
                const long long b0 = (long long)(q / p.q2) * p.s_outer + (long long)(q % p.q2) * p.s_inner;
                const int nu = min(p.U, hid - rank * p.U);
                for (int i = tid; i < (p.T - l) * max(nu, 0); i += 256) {
                    const int t = l + i / nu, uu = rank * p.U + i % nu;
                    p.out[(size_t)(b0 + (long long)t * p.step) * OC + dir * hid + uu] = 0.f;
                }
            }
        }
    }
    // hand-off protocol (CS > 1): every owner sends its h_t to all CTAs of the cluster with st.async; the destination's
    // mbarrier of that buffer counts bytes.  Buffer b is filled during steps s with (s+1)&1 == b and read during step s+1.
    const uint32_t tx_bytes = (uint32_t)nvalid * (uint32_t)hid * 4u;
    const uint32_t mbar0 = cvta_smem(&sm.mbar[0]), mbar1 = cvta_smem(&sm.mbar[1]);
    uint32_t rbase[CS];
    if (CS > 1) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // buffer 1 is filled during step 0
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar1), "r"(tx_bytes) : "memory");
        }
#pragma unroll
        for (int r = 0; r < CS; ++r) rbase[r] = mapa_u32(cvta_smem(&sm), (uint32_t)r);
        cluster.sync();
    } else __syncthreads();

    float4 gxc[NOWN], gxn[NOWN];
    auto load_gx = [&](int s, float4 *dst) {
#pragma unroll
        for (int j = 0; j < NOWN; ++j) {
            if (uvalid && qvalid[j] && s < len[j]) {
                const int t = dir ? len[j] - 1 - s : s;
                dst[j] = __ldg(reinterpret_cast<const float4 *>(p.gx + (size_t)(base[j] + (long long)t * p.step) * GC + (size_t)dir * 4 * hid + (size_t)u * 4));
            } else dst[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    load_gx(0, gxc);

    for (int s = 0; s < maxlen; ++s) {
        const int cur = s & 1, nxt = cur ^ 1;
        if (CS > 1) {
            if (s > 0) mbar_wait_parity(cur ? mbar1 : mbar0, (uint32_t)(((s - 1) >> 1) & 1));
            // every thread of the CTA has now seen this phase complete: without this a warp that owns no cells (padding
            // units) could fall two phases behind and alias the parity bit
            __syncthreads();
            // this buffer's next fill happens during step s+1; nobody can send it before receiving our step-s data
            if (tid == 0 && s + 2 < maxlen)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cur ? mbar1 : mbar0), "r"(tx_bytes) : "memory");
        }
        if (s + 1 < maxlen) load_gx(s + 1, gxn);
        float v[BLT * 4];
#pragma unroll
        for (int b = 0; b < BLT; ++b) {
            const float *hrow = &sm.h[cur][(BLT == 10 ? 0 : ls * 8) + b][ks * 36];
            unsigned long long a01 = 0ull, a23 = 0ull;          // (i,f) and (g,o) partial sums of line b
#pragma unroll
            for (int kq = 0; kq < 8; ++kq) {
                const float4 hv = *reinterpret_cast<const float4 *>(hrow + kq * 4);
                ffma2_bcast(a01, w2[0][kq * 4 + 0], hv.x); ffma2_bcast(a23, w2[1][kq * 4 + 0], hv.x);
                ffma2_bcast(a01, w2[0][kq * 4 + 1], hv.y); ffma2_bcast(a23, w2[1][kq * 4 + 1], hv.y);
                ffma2_bcast(a01, w2[0][kq * 4 + 2], hv.z); ffma2_bcast(a23, w2[1][kq * 4 + 2], hv.z);
                ffma2_bcast(a01, w2[0][kq * 4 + 3], hv.w); ffma2_bcast(a23, w2[1][kq * 4 + 3], hv.w);
            }
            unpack2f(a01, v[b * 4 + 0], v[b * 4 + 1]);
            unpack2f(a23, v[b * 4 + 2], v[b * 4 + 3]);
        }
        // lines 8,9 (BLT == 10): butterfly all-reduce over the 8 k-slices, lanes 0/1 keep line 8/9
        float x89[4] = {0.f, 0.f, 0.f, 0.f};
        if (BLT == 10) {
#pragma unroll
            for (int i = 32; i < 40; ++i) {
                float t = v[i];
                t += __shfl_xor_sync(0xffffffffu, t, 4); t += __shfl_xor_sync(0xffffffffu, t, 2); t += __shfl_xor_sync(0xffffffffu, t, 1);
                v[i] = t;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) x89[g] = (ks & 1) ? v[36 + g] : v[32 + g];
        }
        // lines 0..7: reduce-scatter across the KSPLIT k-slices (adjacent lanes): lane ks ends with lines [ks*LPT, +LPT)
        if (KSPLIT >= 8) {
            const bool up = ks & 4;
#pragma unroll
            for (int i = 0; i < 16; ++i) { const float snd = up ? v[i] : v[i + 16], kp = up ? v[i + 16] : v[i]; v[i] = kp + __shfl_xor_sync(0xffffffffu, snd, 4); }
        }
        if (KSPLIT >= 4) {
            constexpr int n = KSPLIT >= 8 ? 16 : 32;
            const bool up = ks & 2;
#pragma unroll
            for (int i = 0; i < n / 2; ++i) { const float snd = up ? v[i] : v[i + n / 2], kp = up ? v[i + n / 2] : v[i]; v[i] = kp + __shfl_xor_sync(0xffffffffu, snd, 2); }
        }
        if (KSPLIT >= 2) {
            constexpr int n = KSPLIT >= 8 ? 8 : (KSPLIT >= 4 ? 16 : 32);
            const bool up = ks & 1;
#pragma unroll
            for (int i = 0; i < n / 2; ++i) { const float snd = up ? v[i] : v[i + n / 2], kp = up ? v[i + n / 2] : v[i]; v[i] = kp + __shfl_xor_sync(0xffffffffu, snd, 1); }
        }
        // gate non-linearities + state update for the (unit, line) cells this lane owns, then the hand-off
        float hout[NOWN]; bool act[NOWN];
#pragma unroll
        for (int j = 0; j < NOWN; ++j) {
            act[j] = uvalid && qvalid[j] && s < len[j];
            const float pi = (BLT == 10 && j == 1) ? x89[0] : v[j * 4 + 0], pf = (BLT == 10 && j == 1) ? x89[1] : v[j * 4 + 1];
            const float pg = (BLT == 10 && j == 1) ? x89[2] : v[j * 4 + 2], po = (BLT == 10 && j == 1) ? x89[3] : v[j * 4 + 3];
            if (act[j]) {
                const float ig = sigmoidf_acc(pi + gxc[j].x), fg = sigmoidf_acc(pf + gxc[j].y);
                const float gg = tanhf(pg + gxc[j].z), og = sigmoidf_acc(po + gxc[j].w);
                cst[j] = fg * cst[j] + ig * gg;
                hlast[j] = og * tanhf(cst[j]);
            }
            hout[j] = hlast[j];
        }
        if (s + 1 < maxlen) {
#pragma unroll
            for (int j = 0; j < NOWN; ++j) {
                if (uvalid && qvalid[j]) {          // finished lines keep sending their last h so byte counts stay constant
                    const uint32_t off = (uint32_t)((((nxt * BL) + lline[j]) * HLD + (u >> 5) * 36 + (u & 31)) * 4);
                    if (CS > 1) {
                        const uint32_t mb_off = (uint32_t)(sizeof(float) * 2 * BL * HLD) + (uint32_t)nxt * 8u;
#pragma unroll
                        for (int r = 0; r < CS; ++r) st_async_f32(rbase[r] + off, hout[j], rbase[r] + mb_off);
                    } else sm.h[nxt][lline[j]][(u >> 5) * 36 + (u & 31)] = hout[j];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NOWN; ++j) {
            if (act[j]) {
                const int t = dir ? len[j] - 1 - s : s;
                p.out[(size_t)(base[j] + (long long)t * p.step) * OC + dir * hid + u] = hout[j];
            }
            gxc[j] = gxn[j];
        }
        if (CS == 1) __syncthreads();
    }
    if (CS > 1) cluster.sync();          // nobody exits while a peer may still address its shared memory
}

// =============================================================================================
// softmax statistics + arg-max per time step, then CTC best-path collapse
//   (logits / T).softmax(1) ; seq[..., :len].max(dim=0) ; groupby ; drop blank 0
//   rpred.py:226, models.py:115, ctc_decoder.py:63-71
// =============================================================================================
// logits rows [rows][C] (NHWC with H == 1).  One warp per row.
__global__ void k_row_argmax_softmax(const float *__restrict__ logits, long long rows, int C, float temperature,
                                     int *__restrict__ lab, float *__restrict__ conf) {
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float *r = logits + (size_t)row * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, __fdiv_rn(__ldg(r + c), temperature));
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f, be = -1.f; int bi = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
        const float e = expf(__fdiv_rn(__ldg(r + c), temperature) - m);
        s += e;
        if (e > be) { be = e; bi = c; }          // strict >: first maximum within the lane's ascending indices
    }
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        const float oe = __shfl_xor_sync(0xffffffffu, be, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oe > be || (oe == be && oi < bi)) { be = oe; bi = oi; }   // first index on ties, as torch.max
    }
    if (lane == 0) { lab[row] = bi; conf[row] = be / s; }
}

// probs (N, C, W) class-major, as handed to the reference's decoder hook.  One thread per (n, t).
__global__ void k_col_argmax(const float *__restrict__ probs, int N, int C, int W, int *__restrict__ lab, float *__restrict__ conf) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)N * W) return;
    const int t = (int)(idx % W); const long long n = idx / W;
    const float *p = probs + (size_t)n * C * W + t;
    float be = __ldg(p); int bi = 0;
    for (int c = 1; c < C; ++c) { const float v = __ldg(p + (size_t)c * W); if (v > be || (v != v && be == be)) { be = v; bi = c; } }
    lab[idx] = bi; conf[idx] = be;
}

// CTC collapse (ctc_decoder.py:55-72): runs of equal labels; blank (0) runs dropped; (label, first t, last t, max conf of run).
// One block (8 warps) per line: the line's labels and confidences are staged in shared memory (one coalesced pass), pass 1 counts
// the run starts of every 32-step chunk, warp 0 scans the counts, pass 2 writes the runs at their final positions; the walk to a
// run's end reads shared memory.  (History: one warp walking the chunks serially = 7 dependent global round trips, 28 us for
// cfg2's T = 200; then block-parallel with the run walk on global memory = one dependent L2 round trip per time step of the
// longest run, 23 us.)  `staged` = 0 keeps everything in global memory (lines too long for shared memory).
// Optional record assembly at the write: the label becomes its code point (1:1 codec table, kraken/lib/codec.py:164-172;
// 0 = not in the codec) and start / end become positions in the original line image exactly as `_scale_val` computes them
// (kraken/lib/vgsl/rpred.py:138-157,231): int(round(min(max((v * net_scale - padding) * in_scale, 0), width - 1))) with Python's double
// arithmetic - separately rounded multiply / subtract / multiply (no FMA contraction) and round-half-to-even.
struct RecordXform {
    const unsigned *lut; int n_lut;        // label -> code point
    const double *scale;                   // [N][2] = (net_scale, in_scale) per line
    const int *maxv;                       // [N] original line width
    int padding;
};
__device__ __forceinline__ int scale_val_dev(int v, double net_scale, double in_scale, int padding, int maxv) {
    double x = __dmul_rn(__dsub_rn(__dmul_rn((double)v, net_scale), (double)padding), in_scale);
    x = x > 0.0 ? x : 0.0;                                                 // max(x, min_val = 0)
    const double hi = (double)(maxv - 1);
    x = x < hi ? x : hi;                                                   // min(x, max_val - 1)
    return (int)rint(x);
}
__global__ void __launch_bounds__(256) k_ctc_collapse(const int *__restrict__ lab, const float *__restrict__ conf, const int *__restrict__ lens,
                               int N, int T, int max_out, int *__restrict__ o_lab, int *__restrict__ o_start,
                               int *__restrict__ o_end, float *__restrict__ o_conf, int *__restrict__ o_cnt, int staged, RecordXform rx) {
    extern __shared__ int cc_sm[];                    // [chunks + 1]: run starts per chunk, then their exclusive prefix sums; [T] labels; [T] confs
    const int n = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int L = lens ? min(max(lens[n], 0), T) : T;
    const int nch = (L + 31) >> 5;
    const int *l = lab + (size_t)n * T; const float *cf = conf + (size_t)n * T;
    if (staged) {
        int *sl = cc_sm + ((T + 31) / 32 + 1); float *sc = reinterpret_cast<float *>(sl + T);
        for (int t = threadIdx.x; t < L; t += blockDim.x) { sl[t] = __ldg(l + t); sc[t] = __ldg(cf + t); }
        l = sl; cf = sc;
        __syncthreads();
    }
    for (int c = warp; c < nch; c += nw) {
        const int t = 32 * c + lane;
        bool emit = false;
        if (t < L) { const int cur = l[t]; emit = cur != 0 && (t == 0 || l[t - 1] != cur); }
        const unsigned bal = __ballot_sync(0xffffffffu, emit);
        if (lane == 0) cc_sm[c] = __popc(bal);
    }
    __syncthreads();
    if (warp == 0) {
        int carry = 0;
        for (int c0 = 0; c0 < nch; c0 += 32) {
            const int c = c0 + lane;
            const int v = c < nch ? cc_sm[c] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
            if (c < nch) cc_sm[c] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) o_cnt[n] = carry;
    }
    __syncthreads();
    for (int c = warp; c < nch; c += nw) {
        const int t = 32 * c + lane;
        int cur = -1; bool emit = false;
        if (t < L) { cur = l[t]; emit = cur != 0 && (t == 0 || l[t - 1] != cur); }
        const unsigned bal = __ballot_sync(0xffffffffu, emit);
        if (emit) {
            const int pos = cc_sm[c] + __popc(bal & ((1u << lane) - 1));
            int e = t; float mx = cf[t];
            while (e + 1 < L && l[e + 1] == cur) { ++e; mx = fmaxf(mx, cf[e]); }
            if (pos < max_out) {
                int vl = cur, vs = t, ve = e;
                if (rx.scale) {
                    vl = (cur >= 0 && cur < rx.n_lut) ? (int)rx.lut[cur] : 0;
                    const double ns = rx.scale[2 * n], is = rx.scale[2 * n + 1];
                    vs = scale_val_dev(t, ns, is, rx.padding, rx.maxv[n]); ve = scale_val_dev(e, ns, is, rx.padding, rx.maxv[n]);
                }
                o_lab[(size_t)n * max_out + pos] = vl; o_start[(size_t)n * max_out + pos] = vs;
                o_end[(size_t)n * max_out + pos] = ve; o_conf[(size_t)n * max_out + pos] = mx;
            }
        }
    }
}

// probabilities in the reference's (N, C, T) layout (`self.outputs`, rpred.py:227) from NHWC logits rows.
// block = 32 time steps of one line; warp w handles rows w, w+8, ...; smem transpose for coalesced stores.
__global__ void k_probs_nct(const float *__restrict__ logits, float *__restrict__ probs, int T, int C, float temperature) {
    extern __shared__ float pr_sm[];                 // [32][C + 1]
    const int n = blockIdx.y, t0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int ld = C + 1;
    for (int r = wid; r < 32; r += nw) {
        const int t = t0 + r;
        if (t >= T) continue;
        const float *row = logits + ((size_t)n * T + t) * C;
        float m = -INFINITY;
        for (int c = lane; c < C; c += 32) m = fmaxf(m, __fdiv_rn(__ldg(row + c), temperature));
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int c = lane; c < C; c += 32) { const float e = expf(__fdiv_rn(__ldg(row + c), temperature) - m); pr_sm[r * ld + c] = e; s += e; }
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        for (int c = lane; c < C; c += 32) pr_sm[r * ld + c] = pr_sm[r * ld + c] / s;
    }
    __syncthreads();
    for (int c = wid; c < C; c += nw) {
        const int t = t0 + lane;
        if (t < T) probs[((size_t)n * C + c) * T + t] = pr_sm[lane * ld + c];
    }
}

// fallback for class counts whose 32-row tile does not fit shared memory: one warp per (n, t), strided stores
__global__ void k_probs_nct_simple(const float *__restrict__ logits, float *__restrict__ probs, int N, int T, int C, float temperature) {
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (long long)N * T) return;
    const int lane = threadIdx.x & 31;
    const int t = (int)(row % T); const long long n = row / T;
    const float *r = logits + (size_t)row * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, __fdiv_rn(__ldg(r + c), temperature));
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(__fdiv_rn(__ldg(r + c), temperature) - m);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    for (int c = lane; c < C; c += 32) probs[((size_t)n * C + c) * T + t] = expf(__fdiv_rn(__ldg(r + c), temperature) - m) / s;
}

// =============================================================================================
// blla post-net: F.interpolate(o, size) (nearest) -> sigmoid, NHWC logits -> NCHW heat map
// (spred.py:271-272, blla.py:124-125).  src = min(int(floorf(dst * (float)in / out)), in - 1) as ATen.
// =============================================================================================
__global__ void k_upsample_sigmoid(const float *__restrict__ x, float *__restrict__ y, int N, int H, int W, int C, int OH, int OW) {
    const float sh = (float)H / (float)OH, sw = (float)W / (float)OW;
    const long long total = (long long)N * OH * OW;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(idx % OW); long long t = idx / OW;
        const int oy = (int)(t % OH); const long long n = t / OH;
        const int sy = min((int)floorf(oy * sh), H - 1), sx = min((int)floorf(ox * sw), W - 1);
        const float *src = x + ((n * H + sy) * (long long)W + sx) * C;
        for (int c = 0; c < C; ++c) y[((n * C + c) * (long long)OH + oy) * OW + ox] = sigmoidf_acc(__ldg(src + c));
    }
}

}  // namespace kb
