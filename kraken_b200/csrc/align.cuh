// Forced alignment of known label sequences against the recogniser's output, on the device (the `return_logits`
// consumer).  Reference: kraken/tasks/align.py:111-137 (`ForcedAlignmentTaskModel.predict`) with its helpers
//     emission = record.logits.squeeze().log_softmax(0).T      :119      (record.logits = the softmax PROBABILITIES (C, T) of
//                                                                         kraken/lib/vgsl/rpred.py:226-227,200; the second softmax is the reference's)
//     get_trellis   :170-191     backtrack   :194-229     merge_repeats   :232-249
// One CTA per line.  The trellis recursion is sequential in time and parallel over the tokens: one row per __syncthreads, the two live
// rows in shared memory, every row also written to global memory for the backtrack.  Arithmetic follows the reference operation for
// operation: float32 adds / maximum in the recursion, the column-0 prefix sums accumulated in double and rounded per prefix (what
// torch.cumsum does for float32 on the CPU), scores as means of float32 frame probabilities accumulated in double (`sum(...) / n` over
// Python floats).  Only the three transcendental calls (exp / log in the log-softmax, exp of the frame emission) are not torch's own
// code, so scores agree to ~1e-7 relative rather than bit for bit; the token / frame indices are integers and are compared exactly.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace kb { namespace fa {

constexpr int ALIGN_TOO_SHORT = -1;      // fewer output frames than 2 * len(labels): the reference emits an empty record (align.py:113-117)
constexpr int ALIGN_FAILED = -2;         // the backtrack ran out of frames: ValueError("Failed to align") (align.py:228)

struct AlignParams {
    const float *probs;                  // (N, C, T) probabilities, the records' "logits"
    const int *lens;                     // [N] valid frames per line (<= T)
    const int *tokens, *tok_off;         // concatenated label sequences, [N + 1] offsets
    float *trellis;                      // workspace [N][(T + 1) * (Jmax + 1)]
    float *fmax, *flse, *fe0;            // workspace [N][T]: per frame max / log-sum-exp of the probabilities, emission of label 0 (blank)
    int *ptok; float *pprob;             // workspace [N][T]: the path (token index, frame probability) per frame
    int N, C, T, Jmax, max_seg;
    int *seg_token, *seg_start, *seg_end; float *seg_score; int *seg_count;      // [N][max_seg], [N]
    const double *scale; const int *maxv; int padding;      // optional `_scale_val` of the segment borders (as RecordXform), scale == nullptr: frames
};

__device__ __forceinline__ int fa_scale_val(int v, double net_scale, double in_scale, int padding, int maxv) {
    double x = __dmul_rn(__dsub_rn(__dmul_rn((double)v, net_scale), (double)padding), in_scale);
    x = x > 0.0 ? x : 0.0;
    const double hi = (double)(maxv - 1);
    x = x < hi ? x : hi;
    return (int)rint(x);
}

__global__ void __launch_bounds__(256) k_forced_align(AlignParams p) {
    extern __shared__ float fa_sm[];     // [2][Jmax + 1] trellis rows, then [8] reduction values + [8] indices
    const int n = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int T = p.T, C = p.C;
    const int Tl = min(max(p.lens[n], 0), T);
    const int j0 = p.tok_off[n], J = p.tok_off[n + 1] - j0;
    const int *tok = p.tokens + j0;
    if (J <= 0 || Tl < 2 * J) {          // J == 0 is rejected by the host (the reference raises IndexError); kept here for safety
        if (tid == 0) p.seg_count[n] = J <= 0 ? 0 : ALIGN_TOO_SHORT;
        return;
    }
    const float *P = p.probs + (size_t)n * C * T;
    float *M = p.fmax + (size_t)n * T, *L = p.flse + (size_t)n * T, *E0 = p.fe0 + (size_t)n * T;
    const int ld = p.Jmax + 1;
    float *tr = p.trellis + (size_t)n * (T + 1) * ld;
    float *row0 = fa_sm, *row1 = fa_sm + ld;
    float *red_v = fa_sm + 2 * ld; int *red_i = (int *)(red_v + 8);

    // 1. log-softmax statistics of every frame: emission(t, c) = (P[c][t] - M[t]) - L[t]
    for (int t = tid; t < Tl; t += nt) {
        float m = -INFINITY;
        for (int c = 0; c < C; ++c) m = fmaxf(m, __ldg(P + (size_t)c * T + t));
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += expf(__ldg(P + (size_t)c * T + t) - m);
        const float l = logf(s);
        M[t] = m; L[t] = l; E0[t] = (__ldg(P + t) - m) - l;
    }
    __syncthreads();

    // 2. row 0 and column 0 of the trellis
    if (tid == 0) {
        tr[0] = 0.f;
        double acc = 0.0;
        for (int t = 0; t < Tl; ++t) { acc += (double)E0[t]; tr[(size_t)(t + 1) * ld] = (float)acc; }
        for (int r = Tl + 1 - J; r <= Tl; ++r) tr[(size_t)r * ld] = INFINITY;            // trellis[-J:, 0] = inf
    }
    for (int j = 1 + tid; j <= J; j += nt) { tr[j] = -INFINITY; row0[j] = -INFINITY; }     // trellis[0, -J:] = -inf
    __syncthreads();
    if (tid == 0) row0[0] = tr[0];
    __syncthreads();

    // 3. recursion: trellis[t + 1, j] = max(trellis[t, j] + e(t, 0), trellis[t, j - 1] + e(t, tok[j - 1]))
    float *cur = row0, *nxt = row1;
    for (int t = 0; t < Tl; ++t) {
        const float e0 = E0[t], m = M[t], l = L[t];
        float *g = tr + (size_t)(t + 1) * ld;
        for (int j = 1 + tid; j <= J; j += nt) {
            const float ek = (__ldg(P + (size_t)__ldg(tok + j - 1) * T + t) - m) - l;
            const float v = fmaxf(cur[j] + e0, cur[j - 1] + ek);
            nxt[j] = v; g[j] = v;
        }
        if (tid == 0) nxt[0] = g[0];
        __syncthreads();
        float *sw = cur; cur = nxt; nxt = sw;
    }

    // 4. t_start = argmax(trellis[:, J]) - the first maximum
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int r = tid; r <= Tl; r += nt) {
        const float v = tr[(size_t)r * ld + J];
        if (v > bv || (v == bv && r < bi)) { bv = v; bi = r; }
    }
    for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { red_v[tid >> 5] = bv; red_i[tid >> 5] = bi; }
    __syncthreads();
    if (tid != 0) return;
    for (int w = 1; w < (nt >> 5); ++w)
        if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) { bv = red_v[w]; bi = red_i[w]; }
    if (bi == 0x7fffffff) bi = 0;                                   // every value -inf: torch.argmax returns 0
    const int t_start = bi;

    // 5. backtrack (one thread: every step depends on the decision of the one before)
    int *ptok = p.ptok + (size_t)n * T; float *pprob = p.pprob + (size_t)n * T;
    int j = J, t_first = -1;
    for (int t = t_start; t > 0; --t) {
        const float stayed = tr[(size_t)(t - 1) * ld + j] + E0[t - 1];
        const float ek = (__ldg(P + (size_t)tok[j - 1] * T + (t - 1)) - M[t - 1]) - L[t - 1];
        const float changed = tr[(size_t)(t - 1) * ld + j - 1] + ek;
        const bool moved = changed > stayed;
        ptok[t - 1] = j - 1;
        pprob[t - 1] = expf(moved ? ek : E0[t - 1]);
        if (moved && --j == 0) { t_first = t - 1; break; }
    }
    if (t_first < 0) { p.seg_count[n] = ALIGN_FAILED; return; }

    // 6. merge_repeats: runs of one token index -> (token, start, end, mean probability)
    int k = 0, i1 = t_first;
    const double ns = p.scale ? p.scale[2 * n] : 0.0, is = p.scale ? p.scale[2 * n + 1] : 0.0;
    while (i1 < t_start) {
        int i2 = i1; double acc = 0.0;
        const int tk = ptok[i1];
        while (i2 < t_start && ptok[i2] == tk) { acc += (double)pprob[i2]; ++i2; }
        if (k < p.max_seg) {
            const size_t o = (size_t)n * p.max_seg + k;
            p.seg_token[o] = tk;
            p.seg_start[o] = p.scale ? fa_scale_val(i1, ns, is, p.padding, p.maxv[n]) : i1;
            p.seg_end[o] = p.scale ? fa_scale_val(i2, ns, is, p.padding, p.maxv[n]) : i2;
            p.seg_score[o] = (float)(acc / (double)(i2 - i1));
        }
        ++k; i1 = i2;
    }
    p.seg_count[n] = k;
}

inline size_t align_smem(int jmax) { return (size_t)(2 * (jmax + 1) + 16) * sizeof(float); }

} }  // namespace kb::fa
