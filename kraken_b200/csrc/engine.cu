// engine.cu - model object, weight repacking, layer executor and the C ABI of libkraken_b200.so.
// See include/kraken_b200.h for the contract and the reference interfaces each entry point replaces.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <chrono>

#include "../../include/kraken_b200.h"
#include "kernels.cuh"
#include "gemm_tc.cuh"
#include "conv_tc.cuh"
#include "lstm_tc.cuh"
#include "line_prep.cuh"
#include "conv1_tc.cuh"
#include "align.cuh"
#include "vgsl_plan.hpp"

namespace kb {

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
#define CK(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t _e = (call);                                                                                   \
        if (_e != cudaSuccess)                                                                                     \
            throw CudaError(std::string(#call) + " failed: " + cudaGetErrorString(_e) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

struct Tensor {
    float *p = nullptr; int64_t n = 0, h = 0, w = 0, c = 0;
    __half *hi = nullptr, *lo = nullptr;     // optional fp16 operand planes (x1, x2s) written by a fused producer (same layout as p)
    // space-to-depth operand planes [N][ceil(H/2)][ceil(W/2)][s2d_c] (channel = (h&1, w&1, c), zero padded to s2d_c) for a stride-2
    // convolution that runs as a stride-1 convolution on k_conv_tc
    __half *s_hi = nullptr, *s_lo = nullptr; int s2d_c = 0;
    const float *nchw = nullptr;             // network input still in the caller's NCHW layout (first layer consumes it directly)
    int64_t numel() const { return n * h * w * c; }
};

struct LeafWeights {
    std::vector<std::vector<float>> host;      // by slot, reference layout
    std::vector<bool> loaded;
    // device (engine layout)
    float *wt = nullptr;    // [K][Ncp]  conv/linear/lstm-x projection
    float *bias = nullptr;  // [Cout] / folded LSTM bias / GN beta
    float *aux = nullptr;   // W_hh [dirs][4h][h] / GN gamma
    __half *b_hi = nullptr, *b_lo = nullptr;  // [ncols][K] fp16 operand planes for the tcgen05 GEMM (K-major)
    __half *c_hi = nullptr, *c_lo = nullptr;  // [tap][chunk][Cout][32] fp16 operand planes for the tcgen05 convolution
    int s2d = 0, s_th = 0, s_tw = 0, s_py = 0, s_px = 0, s_cs = 0;   // stride-2 conv as space-to-depth stride-1 conv: block taps, block padding, channels
    void *wpk = nullptr;                      // W_hh as fp16 operand planes of the tcgen05 recurrences
    __half *c1_pk = nullptr;                  // first-layer 3x3 conv on tcgen05 (conv1_tc.cuh): [b1 rows | b2s rows] x K16 operand
    float *peep = nullptr;                    // ocropy cell: peephole vectors [dir][ip, fp, op][h]
    std::vector<std::vector<float>> legacy_host;   // legacy cells rewritten as (W_ih, W_hh, b_ih, b_hh) per direction
    float *whh_t = nullptr; int whh_ncp = 0;  // hidden > 256: W_hh^T [dir][k = h][gate column (unit, gate) padded to 64] for the per-step GEMM
    int ncp = 0, K = 0, ncols = 0;
};

struct Arena {
    char *base = nullptr; size_t cap = 0, off = 0; bool dry = false;
    void *alloc(size_t bytes) {
        size_t a = (off + 255) & ~size_t(255);
        off = a + bytes;
        if (dry) return reinterpret_cast<void *>(uintptr_t(256) + a);     // fake, never dereferenced
        if (off > cap) throw CudaError("internal: arena overflow");
        return base + a;
    }
};

struct Lens { bool has = false; std::vector<int32_t> v; };

// Per-call state.  Workspace 0 serves the synchronous entry points on the caller's stream; workspaces 1..depth are the slots of the
// asynchronous pipeline (kb_recognize_async / kb_wait): own stream, arena, pinned result block and range flag each, ONE copy of the
// weights shared by all of them.
struct Workspace {
    int index = 0;
    Arena arena;
    cudaStream_t stream = nullptr;            // internal stream of an async slot (workspace 0 runs on the caller's stream)
    cudaEvent_t done = nullptr, input_ready = nullptr;
    void *pinned = nullptr; size_t pinned_cap = 0;          // result block (D2H target)
    char *params = nullptr; size_t params_cap = 0, params_off = 0;      // pinned staging of the call's small host arrays (lens, widths)
    int *d_flag = nullptr;           // device: set by plane producers when an activation leaves the fp16 range
    int *h_flag = nullptr;           // pinned copy, valid after the stream is synchronised
    uint8_t *u8_raw = nullptr; size_t u8_raw_cap = 0;       // kb_recognize_u8: device copy of host uint8 lines
    float *u8_f32 = nullptr; size_t u8_f32_cap = 0;         // ... and their float form (the network input)
    std::map<std::string, Tensor> taps;
    bool timing = false;
    struct Stage { std::string name; cudaEvent_t a = nullptr, b = nullptr; float ms = 0.f; };
    std::vector<Stage> stages;       // event pool, reused across calls
    size_t n_stages = 0;             // entries used by the most recent call
    // ---- pending asynchronous call (kb_recognize_async .. kb_wait)
    bool busy = false; int64_t ticket = -1;
    struct Pending {
        int n = 0, h = 0, w = 0, T = 0, max_out = 0, dtype = 0; float temperature = 1.f;
        const void *device_input = nullptr;      // network input as staged on the device (valid until the slot is reused)
        bool has_widths = false; std::vector<int32_t> widths, olens;
        char *d_result = nullptr; size_t result_bytes = 0;
    } pend;
    void release() {
        if (arena.base) cudaFree(arena.base);
        arena = Arena();
        if (pinned) cudaFreeHost(pinned);
        pinned = nullptr; pinned_cap = 0;
        if (params) cudaFreeHost(params);
        params = nullptr; params_cap = 0;
        if (h_flag) cudaFreeHost(h_flag);
        h_flag = nullptr;
        if (d_flag) cudaFree(d_flag);
        d_flag = nullptr;
        if (u8_raw) cudaFree(u8_raw);
        if (u8_f32) cudaFree(u8_f32);
        u8_raw = nullptr; u8_f32 = nullptr; u8_raw_cap = u8_f32_cap = 0;
        for (auto &e : stages) { if (e.a) cudaEventDestroy(e.a); if (e.b) cudaEventDestroy(e.b); }
        stages.clear(); n_stages = 0;
        if (done) cudaEventDestroy(done);
        if (input_ready) cudaEventDestroy(input_ready);
        done = input_ready = nullptr;
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
        busy = false; ticket = -1;
    }
};

}  // namespace kb

using namespace kb;

struct kb_model {
    std::unique_ptr<Plan> plan;
    std::vector<LeafWeights> lw;
    int device = -1;
    bool finalized = false;
    std::mutex mu;
    std::vector<std::unique_ptr<Workspace>> wss;      // [0] synchronous calls, [1..] asynchronous slots
    int64_t next_ticket = 0;
    std::vector<void *> dev_allocs;              // weights
    int64_t launches = 0;
    bool timing = false;
    int sm_count = 148;
    int max_clusters8 = -1;          // co-resident 8-CTA clusters of the recurrence kernel (queried once)
    int max_clusters_tc = -1;        // ... of the tcgen05 recurrence
    int fuse_mask = 7;               // bit 0: stencil+pool group, bit 1: tcgen05 conv group, bit 2: stride-2 conv via space-to-depth
    bool keep_fp32 = false;          // KB_KEEP_FP32=1: fused producers also write the fp32 tensor their plane-only consumer ignores (taps)
    bool force_ffma = false;         // second attempt of a call whose first attempt raised the flag: fp32 CUDA-core kernels only
    int64_t overflow_reruns = 0;
    double prof_us[4] = {0, 0, 0, 0}; int64_t prof_calls = 0;    // KB_HOST_PROF: host microseconds in plan / launch / wait / unpack
    bool fuse = true;                // fused layer groups (KB_FUSE=0 runs every layer on its own, e.g. for layer taps)
    bool use_tc = true;              // tcgen05 GEMM path (KB_GEMM=ffma forces the CUDA-core kernel)
    // kb_prepare_lines_u8 scratch: device [page copy | line table | coefficient tables | horizontal-pass rows | line maxima], pinned host mirror of the tables
    unsigned *codec_lut = nullptr; int codec_n = 0;      // kb_model_set_codec: label -> code point of a 1:1 codec (device)
    std::vector<unsigned> codec_host;
    char *prep_dev = nullptr; size_t prep_dev_cap = 0; char *prep_host = nullptr; size_t prep_host_cap = 0;
    kb_model() { wss.emplace_back(new Workspace()); }
    Workspace *ws0() { return wss[0].get(); }
    void release_device_state() {
        for (void *p : dev_allocs) cudaFree(p);
        dev_allocs.clear();
        if (prep_dev) cudaFree(prep_dev);
        if (codec_lut) cudaFree(codec_lut);
        codec_lut = nullptr; codec_n = 0;
        if (prep_host) cudaFreeHost(prep_host);
        prep_dev = prep_host = nullptr; prep_dev_cap = prep_host_cap = 0;
        for (auto &w : wss) w->release();
        wss.resize(1);
    }
    ~kb_model() {
        if (device >= 0) {
            int prev = -1;
            if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
            cudaSetDevice(device);
            release_device_state();
            if (prev >= 0) cudaSetDevice(prev);
        }
    }
};

namespace kb {

#define LAUNCH(m, kern, grid, block, smem, st, ...)                                 \
    do {                                                                            \
        kern<<<grid, block, smem, st>>>(__VA_ARGS__);                               \
        ++(m)->launches;                                                            \
        CK(cudaPeekAtLastError());                                                  \
    } while (0)

// RAII device timer for one named stage of a call (only when kb_set_timing(m, 1)); events are recorded on the
// launching stream, so the elapsed time is what the stream spent between the two records.
struct StageTimer {
    Workspace *ws; cudaStream_t st; int idx = -1;
    StageTimer(Workspace *ws_, cudaStream_t st_, const std::string &name, bool active) : ws(ws_), st(st_) {
        if (!active || !ws->timing) return;
        if (ws->n_stages == ws->stages.size()) {
            Workspace::Stage s; CK(cudaEventCreate(&s.a)); CK(cudaEventCreate(&s.b)); ws->stages.push_back(s);
        }
        idx = (int)ws->n_stages++;
        ws->stages[idx].name = name; ws->stages[idx].ms = 0.f;
        CK(cudaEventRecord(ws->stages[idx].a, st));
    }
    ~StageTimer() { if (idx >= 0) cudaEventRecord(ws->stages[idx].b, st); }
};

static inline unsigned grid1d(long long total, int block, int sm_count) {
    long long g = (total + block - 1) / block;
    long long cap = (long long)sm_count * 16;
    return (unsigned)std::max<long long>(1, std::min(g, cap));
}

// -------------------------------------------------------------------------------------------------
// weights
// -------------------------------------------------------------------------------------------------
static float *upload(kb_model *m, const std::vector<float> &h) {
    float *d = nullptr;
    CK(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(float)));
    m->dev_allocs.push_back(d);
    if (!h.empty()) CK(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    return d;
}

// host-side split of weights into the two fp16 operand planes (kernels.cuh split_f16); false if a value leaves the fp16 range
static bool split_planes_host(const std::vector<float> &v, std::vector<__half> &hi, std::vector<__half> &lo) {
    hi.resize(v.size()); lo.resize(v.size());
    for (size_t i = 0; i < v.size(); ++i) {
        if (!(std::fabs(v[i]) <= 65504.f)) return false;
        const __half h1 = __float2half_rn(v[i]);
        hi[i] = h1; lo[i] = __float2half_rn((v[i] - __half2float(h1)) * X2_SCALE);
    }
    return true;
}
static __half *upload_half(kb_model *m, const std::vector<__half> &h) {
    __half *d = nullptr;
    CK(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(__half)));
    m->dev_allocs.push_back(d);
    if (!h.empty()) CK(cudaMemcpy(d, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice));
    return d;
}
// rows[N][K] (K contiguous) -> device operand planes; a layer whose weights do not fit fp16 keeps the CUDA-core kernel
static void upload_split(kb_model *m, const std::vector<float> &rows, LeafWeights &w) {
    std::vector<__half> hi, lo;
    if (!split_planes_host(rows, hi, lo)) { w.b_hi = w.b_lo = nullptr; return; }
    w.b_hi = upload_half(m, hi); w.b_lo = upload_half(m, lo);
}

// opt-in dynamic shared memory sizes of the tensor-core kernels; function attributes are per device, so this runs at every
// kb_model_finalize (after cudaSetDevice) instead of behind process-wide "already done" flags
static void set_kernel_attributes() {
    CK(cudaFuncSetAttribute(tc::k_gemm_tc<128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GemmCfg<128>::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::k_gemm_tc<256, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GemmCfg<256>::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::k_gemm_tc<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GemmCfg<256>::SMEM_BYTES));
    CK(cudaFuncSetAttribute((tc::k_gemm_tc<256, 0, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GemmCfg<256>::SMEM_BYTES));
    CK(cudaFuncSetAttribute(ctc::k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(c1tc::k_conv1_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c1tc::conv1_tc_smem(32)));
    CK(cudaFuncSetAttribute(c1tc::k_conv1_tc<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c1tc::conv1_tc_smem(16)));
    CK(cudaFuncSetAttribute((ltc::k_lstm_rec_tc<8, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (ltc::ClusterCfg<8, 2>::SMEM_BYTES)));
    CK(cudaFuncSetAttribute((ltc::k_lstm_rec_tc<8, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, (ltc::ClusterCfg<8, 3>::SMEM_BYTES)));
    CK(cudaFuncSetAttribute((ltc::k_lstm_rec_tc<8, 4, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, (ltc::ClusterCfg<8, 4, 1>::SMEM_BYTES)));
    CK(cudaFuncSetAttribute((ltc::k_lstm_rec_tc<16, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (ltc::ClusterCfg<16, 2>::SMEM_BYTES)));
    CK(cudaFuncSetAttribute(ltc::k_lstm_rec_tc_small<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ltc::SmallCfg<64>::SMEM_BYTES));
    CK(cudaFuncSetAttribute(ltc::k_lstm_rec_tc_small<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ltc::SmallCfg<32>::SMEM_BYTES));
    CK(cudaFuncSetAttribute(ltc::k_lstm_rec_tc_small<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ltc::SmallCfg<16>::SMEM_BYTES));
}

static void finalize_weights(kb_model *m) {
    for (void *p : m->dev_allocs) cudaFree(p);
    m->dev_allocs.clear();
    for (size_t li = 0; li < m->plan->leaf_nodes.size(); ++li) {
        const Node &n = *m->plan->leaf_nodes[li];
        LeafWeights &w = m->lw[li];
        w.wt = w.bias = w.aux = nullptr; w.b_hi = w.b_lo = w.c_hi = w.c_lo = nullptr; w.wpk = nullptr; w.whh_t = nullptr; w.c1_pk = nullptr;
        auto need = [&](int slots) {
            for (int s = 0; s < slots; ++s)
                if ((int)w.loaded.size() <= s || !w.loaded[s]) throw SpecError("weights of layer " + n.name + " not loaded (missing tensor for nn." + n.path + ")");
        };
        if (n.kind == K_CONV) {
            need(2);
            const int K = n.kh * n.kw * n.cin, ncp = (n.cout + 63) / 64 * 64;
            std::vector<float> wt((size_t)K * ncp, 0.f);
            const std::vector<float> &src = w.host[0];       // [Cout][Cin][kh][kw]
            for (int co = 0; co < n.cout; ++co)
                for (int ci = 0; ci < n.cin; ++ci)
                    for (int ky = 0; ky < n.kh; ++ky)
                        for (int kx = 0; kx < n.kw; ++kx)
                            wt[(size_t)((ky * n.kw + kx) * n.cin + ci) * ncp + co] = src[(((size_t)co * n.cin + ci) * n.kh + ky) * n.kw + kx];
            w.wt = upload(m, wt); w.bias = upload(m, w.host[1]); w.ncp = ncp; w.K = K; w.ncols = n.cout;
            if (n.kh == 1 && n.kw == 1) upload_split(m, src, w);          // [Cout][Cin] is already K-major
            if (n.cin == 1 && n.kh == 3 && n.kw == 3 && n.sy == 1 && n.sx == 1 && n.dy == 1 && n.dx == 1 && (n.cout == 16 || n.cout == 32)) {
                std::vector<__half> pk;
                if (c1tc::pack_conv1_weights(src, n.cout, pk)) w.c1_pk = upload_half(m, pk);
            }
            if (n.cin % 32 == 0 && n.sy == 1 && n.sx == 1 && n.dy == 1 && n.dx == 1 && n.cout % 32 == 0 && n.kh + 1 <= ctc::MAX_ROWS &&
                n.kh * n.kw > 1) {
                // tcgen05 convolution operand: [tap][32-channel chunk][Cout][32] fp16 planes
                const int nc = n.cin / 32;
                std::vector<float> taps((size_t)n.kh * n.kw * nc * n.cout * 32);
                for (int ky = 0; ky < n.kh; ++ky)
                    for (int kx = 0; kx < n.kw; ++kx)
                        for (int cc = 0; cc < nc; ++cc)
                            for (int co = 0; co < n.cout; ++co)
                                for (int ci = 0; ci < 32; ++ci)
                                    taps[((((size_t)(ky * n.kw + kx) * nc + cc) * n.cout) + co) * 32 + ci] =
                                        src[(((size_t)co * n.cin + cc * 32 + ci) * n.kh + ky) * n.kw + kx];
                std::vector<__half> hi, lo;
                if (split_planes_host(taps, hi, lo)) { w.c_hi = upload_half(m, hi); w.c_lo = upload_half(m, lo); }
            }
            w.s2d = 0;
            if (n.sy == 2 && n.sx == 2 && n.dy == 1 && n.dx == 1 && n.cout % 32 == 0 && n.kh * n.kw > 1) {
                // stride 2 = stride 1 over the space-to-depth input: in(2(y+by)+py, 2(x+bx)+px) with ky = 2 by + py + pad.
                // Block taps by in [by0, by1]; weights [tap][32-channel chunk][Cout][32] over channels (py, px, c), zero where
                // (ky, kx) falls outside the filter or the channel is padding.
                auto range = [](int k, int pad, int *b0, int *b1) {
                    *b0 = 99; *b1 = -99;
                    for (int b = -8; b <= 8; ++b) for (int ph = 0; ph < 2; ++ph) { const int kk = 2 * b + ph + pad; if (kk >= 0 && kk < k) { *b0 = std::min(*b0, b); *b1 = std::max(*b1, b); } }
                };
                int by0, by1, bx0, bx1;
                range(n.kh, n.py, &by0, &by1); range(n.kw, n.px, &bx0, &bx1);
                const int th = by1 - by0 + 1, tw = bx1 - bx0 + 1, cs = (4 * n.cin + 31) / 32 * 32, nc = cs / 32;
                if (th + 1 <= ctc::MAX_ROWS && by0 <= 0 && bx0 <= 0 && ctc::conv_tc_plan(th, tw, n.cout, nullptr, nullptr, nullptr) <= 227 * 1024) {
                    std::vector<float> taps((size_t)th * tw * nc * n.cout * 32, 0.f);
                    for (int tby = 0; tby < th; ++tby)
                        for (int tbx = 0; tbx < tw; ++tbx)
                            for (int sc = 0; sc < 4 * n.cin; ++sc) {
                                const int ph = sc / n.cin, ci = sc % n.cin, ky = 2 * (tby + by0) + (ph >> 1) + n.py, kx = 2 * (tbx + bx0) + (ph & 1) + n.px;
                                if (ky < 0 || ky >= n.kh || kx < 0 || kx >= n.kw) continue;
                                for (int co = 0; co < n.cout; ++co)
                                    taps[((((size_t)(tby * tw + tbx) * nc + sc / 32) * n.cout) + co) * 32 + (sc % 32)] =
                                        src[(((size_t)co * n.cin + ci) * n.kh + ky) * n.kw + kx];
                            }
                    std::vector<__half> hi, lo;
                    if (split_planes_host(taps, hi, lo)) {
                        w.c_hi = upload_half(m, hi); w.c_lo = upload_half(m, lo);
                        w.s2d = 1; w.s_th = th; w.s_tw = tw; w.s_py = -by0; w.s_px = -bx0; w.s_cs = cs;
                    }
                }
            }
        } else if (n.kind == K_LINEAR) {
            need(2);
            const int K = n.cin, ncp = (n.cout + 63) / 64 * 64, ld = n.cin + (n.aug ? 1 : 0);
            std::vector<float> wt((size_t)K * ncp, 0.f), b(w.host[1]);
            for (int co = 0; co < n.cout; ++co) {
                for (int ci = 0; ci < n.cin; ++ci) wt[(size_t)ci * ncp + co] = w.host[0][(size_t)co * ld + ci + (n.aug ? 1 : 0)];
                if (n.aug) b[co] += w.host[0][(size_t)co * ld];          // the constant-one input column (layers.py:718-719)
            }
            w.wt = upload(m, wt); w.bias = upload(m, b); w.ncp = ncp; w.K = K; w.ncols = n.cout;
            {
                std::vector<float> rows((size_t)n.cout * K);
                for (int co = 0; co < n.cout; ++co)
                    for (int ci = 0; ci < K; ++ci) rows[(size_t)co * K + ci] = w.host[0][(size_t)co * ld + ci + (n.aug ? 1 : 0)];
                upload_split(m, rows, w);
            }
        } else if (n.kind == K_GN) {
            need(2);
            w.aux = upload(m, w.host[0]); w.bias = upload(m, w.host[1]);
        } else if (n.kind == K_LSTM) {
            const int dirs = n.bidi ? 2 : 1, h = n.hidden, gc = dirs * 4 * h;
            const int K = n.cin, ncp = (gc + 63) / 64 * 64;
            if (n.legacy) {
                // legacy cells -> the standard layout: the constant-one input column of weight_ih IS the bias (layers.py:522-524), no b_hh;
                // the ocropy cell keeps its peephole vectors for the per-step kernel
                if (n.legacy == 2 && !n.bidi) throw SpecError(n.name + ": the ocropy cell is always bidirectional (PeepholeBidiLSTM); the reference's forward fails for a one-directional spec");
                for (int d = 0; d < dirs; ++d)
                    for (int sl = 0; sl < (n.legacy == 2 ? 5 : 2); ++sl)
                        if ((int)w.loaded.size() <= d * 5 + sl || !w.loaded[d * 5 + sl]) throw SpecError("weights of layer " + n.name + " not loaded (missing tensor for nn." + n.path + ")");
                std::vector<std::vector<float>> std4((size_t)4 * dirs);
                std::vector<float> peep;
                for (int d = 0; d < dirs; ++d) {
                    const std::vector<float> &wl = w.host[d * 5 + 0];        // [4h][K + 1]
                    std4[d * 4 + 0].resize((size_t)4 * h * K); std4[d * 4 + 2].resize((size_t)4 * h); std4[d * 4 + 3].assign((size_t)4 * h, 0.f);
                    for (int r = 0; r < 4 * h; ++r) {
                        std4[d * 4 + 2][r] = wl[(size_t)r * (K + 1)];
                        memcpy(&std4[d * 4 + 0][(size_t)r * K], &wl[(size_t)r * (K + 1) + 1], (size_t)K * sizeof(float));
                    }
                    std4[d * 4 + 1] = w.host[d * 5 + 1];
                    if (n.legacy == 2) for (int k = 0; k < 3; ++k) peep.insert(peep.end(), w.host[d * 5 + 2 + k].begin(), w.host[d * 5 + 2 + k].end());
                }
                w.legacy_host = std::move(std4);
                w.peep = n.legacy == 2 ? upload(m, peep) : nullptr;
            } else need(4 * dirs);
            const std::vector<std::vector<float>> &H4 = n.legacy ? w.legacy_host : w.host;
            std::vector<float> wt((size_t)K * ncp, 0.f), b((size_t)gc, 0.f), whh((size_t)dirs * 4 * h * h);
            for (int d = 0; d < dirs; ++d) {
                const std::vector<float> &wih = H4[d * 4 + 0], &wh = H4[d * 4 + 1], &bih = H4[d * 4 + 2], &bhh = H4[d * 4 + 3];
                for (int g = 0; g < 4; ++g)
                    for (int u = 0; u < h; ++u) {
                        const int col = d * 4 * h + u * 4 + g, row = g * h + u;      // torch gate order i,f,g,o
                        for (int ci = 0; ci < K; ++ci) wt[(size_t)ci * ncp + col] = wih[(size_t)row * K + ci];
                        b[col] = bih[row] + bhh[row];
                    }
                std::copy(wh.begin(), wh.end(), whh.begin() + (size_t)d * 4 * h * h);
            }
            w.wt = upload(m, wt); w.bias = upload(m, b); w.aux = upload(m, whh); w.ncp = ncp; w.K = K; w.ncols = gc;
            w.whh_t = nullptr;
            if (h > 256 || n.legacy == 2 || getenv("KB_LSTM_GENERIC")) {
                // generic per-step path: W_hh^T [dir][k][col = 4 u + gate], columns padded to the GEMM's tile
                const int hncp = (4 * h + 63) / 64 * 64;
                std::vector<float> t((size_t)dirs * h * hncp, 0.f);
                for (int d = 0; d < dirs; ++d) {
                    const std::vector<float> &wh = H4[d * 4 + 1];          // [4h][h], torch gate order i,f,g,o
                    for (int g = 0; g < 4; ++g)
                        for (int u = 0; u < h; ++u)
                            for (int k = 0; k < h; ++k) t[((size_t)d * h + k) * hncp + u * 4 + g] = wh[(size_t)(g * h + u) * h + k];
                }
                w.whh_t = upload(m, t); w.whh_ncp = hncp;
            }
            {
                std::vector<float> rows((size_t)gc * K);
                for (int d = 0; d < dirs; ++d)
                    for (int g = 0; g < 4; ++g)
                        for (int u = 0; u < h; ++u)
                            memcpy(&rows[(size_t)(d * 4 * h + u * 4 + g) * K], &H4[d * 4][(size_t)(g * h + u) * K], (size_t)K * sizeof(float));
                upload_split(m, rows, w);
            }
            if (h > 128 && h <= 256) {
                // tcgen05 recurrence operand (lstm_tc.cuh): per (dir, cluster rank) 128 gate rows (row = 4*slot + gate, 32 unit slots per
                // CTA) x K = 256 (k = 32*rank' + slot'); plane 0 = fp16(W), plane 1 = fp16((W - plane0) * 2^11); plain
                // row-major: every thread of the kernel copies its own gate row into tensor memory
                const int U = (h + 7) / 8;
                std::vector<uint16_t> pk((size_t)dirs * 8 * 2 * 128 * 256, 0);      // [dir][rank][plane][gate row][k] fp16, plain row-major
                for (int d = 0; d < dirs; ++d) {
                    const std::vector<float> &wh = H4[d * 4 + 1];      // [4h][h]
                    for (int r = 0; r < 8; ++r)
                        for (int mrow = 0; mrow < 128; ++mrow) {
                            const int slot = mrow >> 2, gate = mrow & 3, u = r * U + slot;
                            if (slot >= U || u >= h) continue;
                            for (int kp = 0; kp < 256; ++kp) {
                                const int r2 = kp >> 5, s2 = kp & 31, u2 = r2 * U + s2;
                                if (s2 >= U || u2 >= h) continue;
                                const float x = wh[(size_t)(gate * h + u) * h + u2];
                                const __half x1 = __float2half_rn(x);
                                const size_t o = ((((size_t)d * 8 + r) * 2) * 128 + mrow) * 256 + kp;
                                pk[o] = __half_as_ushort(x1);
                                pk[o + (size_t)128 * 256] = __half_as_ushort(__float2half_rn((x - __half2float(x1)) * X2_SCALE));
                            }
                        }
                }
                void *dp = nullptr;
                CK(cudaMalloc(&dp, pk.size() * sizeof(uint16_t)));
                m->dev_allocs.push_back(dp);
                CK(cudaMemcpy(dp, pk.data(), pk.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                w.wpk = dp;
            } else if (h <= 32) {
                // single-CTA tcgen05 recurrence (k_lstm_rec_tc_small): [dir][plane][gate row = 4*unit + gate][k = 32] fp16
                std::vector<uint16_t> pk((size_t)dirs * 2 * 128 * 32, 0);
                for (int d = 0; d < dirs; ++d) {
                    const std::vector<float> &wh = H4[d * 4 + 1];      // [4h][h]
                    for (int u = 0; u < h; ++u)
                        for (int gate = 0; gate < 4; ++gate)
                            for (int u2 = 0; u2 < h; ++u2) {
                                const float x = wh[(size_t)(gate * h + u) * h + u2];
                                const __half x1 = __float2half_rn(x);
                                const size_t o = (((size_t)d * 2) * 128 + (size_t)(4 * u + gate)) * 32 + u2;
                                pk[o] = __half_as_ushort(x1);
                                pk[o + (size_t)128 * 32] = __half_as_ushort(__float2half_rn((x - __half2float(x1)) * X2_SCALE));
                            }
                }
                void *dp = nullptr;
                CK(cudaMalloc(&dp, pk.size() * sizeof(uint16_t)));
                m->dev_allocs.push_back(dp);
                CK(cudaMemcpy(dp, pk.data(), pk.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                w.wpk = dp;
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// executor
// -------------------------------------------------------------------------------------------------
struct Exec {
    kb_model *m; Workspace *ws; cudaStream_t st; bool dry;
    int leaf_counter = 0;
    bool planes_hint = false;        // set by run() for a GroupNorm / LSTM whose consumer reads fp16 operand planes
    bool s2d_hint = false;           // ... for a GroupNorm whose consumer is a stride-2 convolution on the tensor cores

    int *dev_lens(const Lens &l) {
        if (!l.has) return nullptr;
        int *d = (int *)ws->arena.alloc(l.v.size() * sizeof(int));
        if (!dry) CK(cudaMemcpyAsync(d, l.v.data(), l.v.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        return d;
    }
    Tensor mk(const Dims &d) {
        Tensor t; t.n = d.n; t.c = d.c; t.h = d.h; t.w = d.w;
        t.p = (float *)ws->arena.alloc((size_t)std::max<int64_t>(t.numel(), 1) * sizeof(float));
        return t;
    }
    static Dims dims_of(const Tensor &t) { Dims d; d.n = t.n; d.c = t.c; d.h = t.h; d.w = t.w; return d; }

    // tcgen05 path: plain per-pixel GEMMs (1x1 stride-1 conv, Linear, LSTM input projection) whose K rows can be TMA'd
    bool tc_eligible(const Tensor &x, const LeafWeights &w, const Node *conv) const {
        if (!m->use_tc || !w.b_hi) return false;
        if (conv && (conv->kh != 1 || conv->kw != 1 || conv->sy != 1 || conv->sx != 1)) return false;
        const long long M = (long long)x.n * x.h * x.w;
        return (w.K % 8) == 0 && w.K >= 32 && w.ncols >= 64 && M >= 128;        // K % 8: 16-byte row pitch of the fp16 planes (TMA)
    }
    // recognition head: when set (kb_recognize without a probability request), the network's final Linear runs with the arg-max
    // epilogue of k_gemm_tc<256, 1> and emits per-row labels / confidences instead of logits
    struct ArgmaxOut { float temperature = 1.f; int *lab = nullptr; float *conf = nullptr; bool done = false; };
    ArgmaxOut *amx = nullptr;
    const Node *final_leaf = nullptr;

    void gemm_tc(const Tensor &x, const LeafWeights &w, int act, float *y, ArgmaxOut *am = nullptr) {
        const long long M = (long long)x.n * x.h * x.w;
        const int K = w.K, N = w.ncols;
        __half *a_hi = x.hi, *a_lo = x.lo;
        if (!a_hi) {
            a_hi = (__half *)ws->arena.alloc((size_t)M * K * sizeof(__half));
            a_lo = (__half *)ws->arena.alloc((size_t)M * K * sizeof(__half));
        }
        if (dry) return;
        if (!x.hi) LAUNCH(m, tc::k_split_f16, grid1d(M * K / 4, 256, m->sm_count), 256, 0, st, x.p, a_hi, a_lo, M * K / 4, ws->d_flag);
        CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
        const uint32_t bkh = tc::BK;      // 64-byte operand rows; 128-byte rows (BKH = 64, two deeper stages) measured 0.1233 vs 0.1209 ms on cfg2's projection
        if (!tc::make_map_2d(&ta_hi, a_hi, (uint64_t)M, (uint64_t)K, tc::BM, bkh) || !tc::make_map_2d(&ta_lo, a_lo, (uint64_t)M, (uint64_t)K, tc::BM, bkh) ||
            !tc::make_map_2d(&tb_hi, w.b_hi, (uint64_t)N, (uint64_t)K, 128, bkh) || !tc::make_map_2d(&tb_lo, w.b_lo, (uint64_t)N, (uint64_t)K, 128, bkh))
            throw CudaError("cuTensorMapEncodeTiled failed");
        tc::GemmTcParams gp; gp.c = y; gp.bias = w.bias; gp.M = (int)M; gp.N = N; gp.K = K; gp.ldc = N; gp.act = act;
        gp.lab = am ? am->lab : nullptr; gp.conf = am ? am->conf : nullptr; gp.temperature = am ? am->temperature : 1.f;
        const int tiles_m = (int)((M + tc::BM - 1) / tc::BM);
        // N <= 128: 128-wide tiles with two accumulator sets (epilogue under the next tile's MMAs).  Wider outputs stay on 256-wide
        // tiles: measured on cfg2's projection (12800 x 2048 x 768) the 128-wide double-buffered variant moves 1.33x the operand bytes
        // per MAC through L2 and ends up L2-bound at the same 0.12 ms the 256-wide tile spends with its serialised epilogue
        // (profiles/r02_*), and it leaves less L2 bandwidth to the kernels of the other batches in flight.  KB_GEMM_BN=128 forces it.
        const bool force128 = getenv("KB_GEMM_BN") && atoi(getenv("KB_GEMM_BN")) == 128;
        const bool wide = am || (N > 128 && !force128);
        // KB_GEMM_MC=1: 2-CTA clusters with the weight tile multicast for wide outputs (a third fewer bytes from L2 per k-block).  Measured
        // on cfg2's projection in both rounds (TF32 planes: 0.241 vs 0.239 ms; fp16 planes: 0.1217 vs 0.1210 ms): no gain - the L2 -> SM
        // feed is not what holds the mainloop at 57 % tensor-pipe activity - so it stays an option, bit-identical to the default.
        const bool mc = !am && wide && tiles_m >= 4 && getenv("KB_GEMM_MC") && atoi(getenv("KB_GEMM_MC")) == 1;
        if (mc) {
            const int npairs = ((tiles_m + 1) / 2) * ((N + 255) / 256);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(2 * std::min(npairs, std::max(1, m->sm_count / 2))), 1, 1);
            cfg.blockDim = dim3(tc::THREADS, 1, 1);
            cfg.dynamicSmemBytes = tc::GemmCfg<256>::SMEM_BYTES; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, tc::k_gemm_tc<256, 0, 2>, ta_hi, ta_lo, tb_hi, tb_lo, gp));
            ++m->launches;
            CK(cudaPeekAtLastError());
        }
        else if (am) LAUNCH(m, (tc::k_gemm_tc<256, 1>), (unsigned)std::min(tiles_m, m->sm_count), tc::THREADS, tc::GemmCfg<256>::SMEM_BYTES, st, ta_hi, ta_lo, tb_hi, tb_lo, gp);
        else if (wide) LAUNCH(m, (tc::k_gemm_tc<256, 0>), (unsigned)std::min(tiles_m * ((N + 255) / 256), m->sm_count), tc::THREADS, tc::GemmCfg<256>::SMEM_BYTES, st, ta_hi, ta_lo, tb_hi, tb_lo, gp);
        else LAUNCH(m, (tc::k_gemm_tc<128, 0>), (unsigned)std::min(tiles_m * ((N + 127) / 128), m->sm_count), tc::THREADS, tc::GemmCfg<128>::SMEM_BYTES, st, ta_hi, ta_lo, tb_hi, tb_lo, gp);
    }

    void gemm(const Tensor &x, const LeafWeights &w, const Node *conv, int act, float *y, int64_t Ho, int64_t Wo) {
        if (tc_eligible(x, w, conv)) { gemm_tc(x, w, act, y); return; }
        if (dry) return;
        ConvParams p;
        p.x = x.p; p.wt = w.wt; p.bias = w.bias; p.y = y;
        p.N = (int)x.n; p.H = (int)x.h; p.W = (int)x.w; p.Cin = (int)x.c; p.Ho = (int)Ho; p.Wo = (int)Wo;
        p.Cout = w.ncols; p.Ncp = w.ncp;
        if (conv) { p.kh = conv->kh; p.kw = conv->kw; p.sy = conv->sy; p.sx = conv->sx; p.dy = conv->dy; p.dx = conv->dx; p.py = conv->py; p.px = conv->px; }
        else { p.kh = p.kw = p.sy = p.sx = p.dy = p.dx = 1; p.py = p.px = 0; }
        p.K = w.K; p.M = (long long)x.n * Ho * Wo; p.act = act;
        if (p.M == 0) return;
        dim3 grid((unsigned)((p.M + CG_BM - 1) / CG_BM), (unsigned)(w.ncp / CG_BN));
        if ((x.c & 3) == 0) LAUNCH(m, k_conv_gemm<true>, grid, CG_NT, 0, st, p);
        else LAUNCH(m, k_conv_gemm<false>, grid, CG_NT, 0, st, p);
    }

    // Recurrence for hidden sizes above 256 (no resident-weight kernel): per time step one fp32 GEMM launch + one pointwise launch
    // (kernels.cuh k_lstm_generic_*).  Slow (2 T launches) but complete: any nn.LSTM the reference builds (layers.py:507-511) runs.
    void run_lstm_generic(const Node &n, const LeafWeights &w, const LstmParams &lp, float *G, float *H, float *C, const Lens &lens) {
        const int hid = lp.hid, nseq = lp.nseq;
        int maxlen = lp.T;
        if (lp.lens && lens.has) { maxlen = 0; for (int32_t l : lens.v) maxlen = std::max(maxlen, std::min(std::max((int)l, 0), lp.T)); }
        if (getenv("KB_DEBUG")) fprintf(stderr, "[kb] %s: generic per-step recurrence (hidden %d), %d steps x %d directions\n", n.name.c_str(), hid, maxlen, lp.dirs);
        LAUNCH(m, k_lstm_generic_init, grid1d((long long)nseq * lp.T * lp.dirs * hid, 256, m->sm_count), 256, 0, st, lp.out, H, C, lp.lens, nseq, lp.T, hid, lp.dirs, lp.q2,
               lp.s_outer, lp.s_inner, lp.step);
        for (int t = 0; t < maxlen; ++t)
            for (int d = 0; d < lp.dirs; ++d) {
                float *Hd = H + (size_t)d * nseq * hid, *Cd = C + (size_t)d * nseq * hid;
                ConvParams p;
                p.x = Hd; p.wt = w.whh_t + (size_t)d * hid * w.whh_ncp; p.bias = nullptr; p.y = G;
                p.N = 1; p.H = 1; p.W = nseq; p.Cin = hid; p.Ho = 1; p.Wo = nseq; p.Cout = 4 * hid; p.Ncp = w.whh_ncp;
                p.kh = p.kw = p.sy = p.sx = p.dy = p.dx = 1; p.py = p.px = 0; p.K = hid; p.M = nseq; p.act = ACT_LINEAR;
                dim3 grid((unsigned)((p.M + CG_BM - 1) / CG_BM), (unsigned)(w.whh_ncp / CG_BN));
                if ((hid & 3) == 0) LAUNCH(m, k_conv_gemm<true>, grid, CG_NT, 0, st, p);
                else LAUNCH(m, k_conv_gemm<false>, grid, CG_NT, 0, st, p);
                LstmStepParams sp;
                sp.G = G; sp.gx = lp.gx; sp.hstate = Hd; sp.cstate = Cd; sp.out = lp.out; sp.lens = lp.lens; sp.nseq = nseq; sp.T = lp.T; sp.hid = hid;
                sp.dirs = lp.dirs; sp.dir = d; sp.t = t; sp.q2 = lp.q2; sp.s_outer = lp.s_outer; sp.s_inner = lp.s_inner; sp.step = lp.step;
                sp.peep = w.peep ? w.peep + (size_t)d * 3 * hid : nullptr;
                LAUNCH(m, k_lstm_generic_step, (unsigned)(((long long)nseq * hid + 255) / 256), 256, 0, st, sp);
            }
    }

    Tensor leaf(const Node &n, const Tensor &x, Lens &lens) {
        const Dims din = dims_of(x);
        const Dims dout = leaf_dims(n, din);
        const LeafWeights &w = m->lw[n.leaf_index];
        Tensor y;
        const int sm = m->sm_count;
        switch (n.kind) {
        case K_CONV: {
            y = mk(dout);
            gemm(x, w, &n, n.act, y.p, dout.h, dout.w);
            if (!dry) {
                if (n.act == ACT_SOFTMAX) {
                    long long rows = (long long)y.n * y.h * y.w;
                    LAUNCH(m, k_softmax_rows, (unsigned)((rows + 7) / 8), 256, 0, st, y.p, rows, (int)y.c);
                }
            }
            break;
        }
        case K_LINEAR: {
            if (amx && &n == final_leaf && x.h == 1 && w.ncols <= 256 && tc_eligible(x, w, nullptr)) {
                // recognition head: labels + confidences straight from the accumulators, the logits tensor never exists
                y.n = dout.n; y.c = dout.c; y.h = dout.h; y.w = dout.w; y.p = nullptr;
                const size_t rows = (size_t)x.n * x.w;
                amx->lab = (int *)ws->arena.alloc(rows * 4); amx->conf = (float *)ws->arena.alloc(rows * 4);
                amx->done = true;
                gemm_tc(x, w, ACT_LINEAR, nullptr, amx);
                break;
            }
            y = mk(dout);
            gemm(x, w, nullptr, ACT_LINEAR, y.p, x.h, x.w);
            break;
        }
        case K_POOL: {
            y = mk(dout);
            if (!dry && y.numel()) {
                long long total = y.numel() / (((int)x.c & 3) == 0 ? 4 : 1);
                LAUNCH(m, k_maxpool, grid1d(total, 256, sm), 256, 0, st, x.p, y.p, (int)x.n, (int)x.h, (int)x.w, (int)x.c,
                       (int)y.h, (int)y.w, n.kh, n.kw, n.sy, n.sx);
            }
            break;
        }
        case K_RESHAPE: {
            y = mk(dout);
            ReshapeParams rp;
            int64_t i4[4] = {din.n, din.c, din.h, din.w}, o4[4], s5[5]; int dest;
            reshape_dims(i4, n, o4, rp.perm, &dest, s5);
            for (int i = 0; i < 4; ++i) { rp.in_dims[i] = i4[i]; rp.out_dims[i] = o4[i]; }
            for (int i = 0; i < 5; ++i) rp.shape5[i] = s5[i];
            rp.src = n.rs_src; rp.dest = dest;
            if (!dry && y.numel()) LAUNCH(m, k_reshape, grid1d(y.numel(), 256, sm), 256, 0, st, x.p, y.p, rp);
            break;
        }
        case K_DROPOUT: case K_IDENTITY: y = x; break;
        case K_ADD: {
            y = mk(dout);
            if (!dry && y.numel()) LAUNCH(m, k_addition, grid1d(y.numel(), 256, sm), 256, 0, st, x.p, y.p, (long long)x.n, (long long)x.c,
                                          (long long)x.h, (long long)x.w, n.add_dim, n.add_chunk);
            break;
        }
        case K_GN: {
            y = mk(dout);
            const int N = (int)x.n, H = (int)x.h, W = (int)x.w, C = (int)x.c, G = n.groups;
            const long long npix = (long long)H * W;
            int chunks = (int)std::min<long long>(std::max<long long>(1, npix * C / 16384), (long long)std::max(1, 8 * sm / std::max(N, 1)));
            const int cthreads = std::min(C, 256), rows = std::max(1, 256 / cthreads), cpt = (C + cthreads - 1) / cthreads;
            double *partial = (double *)ws->arena.alloc((size_t)N * chunks * G * 2 * sizeof(double));
            float2 *stats = (float2 *)ws->arena.alloc((size_t)N * G * sizeof(float2));
            bool ragged = false;
            if (lens.has) for (int32_t l : lens.v) if (l < W) ragged = true;
            int *dl = ragged ? dev_lens(lens) : nullptr;
            const bool vec4 = (C % 4) == 0 && C <= 1024 && !(getenv("KB_GN") && strcmp(getenv("KB_GN"), "scalar") == 0);
            float2 *ab = vec4 ? (float2 *)ws->arena.alloc((size_t)N * C * sizeof(float2)) : nullptr;
            const bool s2d = vec4 && s2d_hint && (C % 8) == 0;
            const bool planes = vec4 && planes_hint && !s2d;
            planes_hint = false; s2d_hint = false;
            if (planes) { y.hi = (__half *)ws->arena.alloc((size_t)y.numel() * 2); y.lo = (__half *)ws->arena.alloc((size_t)y.numel() * 2); }
            if (s2d) {
                const size_t se = (size_t)N * ((H + 1) / 2) * ((W + 1) / 2) * 4 * C;
                y.s_hi = (__half *)ws->arena.alloc(se * 2); y.s_lo = (__half *)ws->arena.alloc(se * 2); y.s2d_c = 4 * C;
            }
            if (!dry && y.numel()) {
                if (vec4) {
                    // float4 streaming passes (kernels.cuh k_gn_stats4 / k_gn_apply4); the apply pass also emits the fp16 operand planes a
                    // tensor-core consumer wants, which saves the separate split pass (and the fp32 copy unless KB_KEEP_FP32)
                    const int rows4 = 256 / (C / 4);
                    LAUNCH(m, k_gn_stats4, dim3(chunks, N), 256, 0, st, x.p, partial, H, W, C, G, dl, chunks);
                    LAUNCH(m, k_gn_finalize, (unsigned)((N * G + 127) / 128), 128, 0, st, partial, stats, N, G, chunks, H, W, C, dl, 1e-5f);
                    LAUNCH(m, k_gn_coeffs, (unsigned)((N * C + 255) / 256), 256, 0, st, stats, w.aux, w.bias, ab, N, C, G);
                    const long long nb = std::min<long long>((npix + rows4 - 1) / rows4, std::max(1, 16 * sm / std::max(N, 1)));
                    LAUNCH(m, k_gn_apply4, dim3((unsigned)nb, N), 256, 0, st, x.p, ((planes || s2d) && !m->keep_fp32) ? nullptr : y.p,
                           s2d ? y.s_hi : y.hi, s2d ? y.s_lo : y.lo, ab, H, W, C, dl, ws->d_flag, s2d ? 1 : 0);
                } else {
                    const int bt = rows * cthreads;
                    LAUNCH(m, k_gn_partial, dim3(chunks, N), bt, (size_t)bt * cpt * 2 * sizeof(double), st, x.p, partial, H, W, C, G, dl, chunks, rows, cthreads, cpt);
                    LAUNCH(m, k_gn_finalize, (unsigned)((N * G + 127) / 128), 128, 0, st, partial, stats, N, G, chunks, H, W, C, dl, 1e-5f);
                    LAUNCH(m, k_gn_apply, grid1d(y.numel(), 256, sm), 256, 0, st, x.p, y.p, stats, w.aux, w.bias, (long long)y.numel(), H, W, C, G, dl);
                }
            }
            break;
        }
        case K_LSTM: {
            const int dirs = n.bidi ? 2 : 1, hid = n.hidden;
            const bool packed = !n.transpose && lens.has;
            if (packed && x.h != 1)
                throw ShapeError("Height has to be 1 (not " + std::to_string(x.h) + ") for batching/multi-sequences.");   // layers.py:529-530
            if (n.summarize && !n.transpose && lens.has) {
                int32_t mx = 0; for (int32_t l : lens.v) mx = std::max(mx, l);
                if (mx > 1) throw ShapeError("Do not use summarizing layer in x-axis with batching/sequences");           // layers.py:545-546
            }
            Dims dg = din; dg.c = dirs * 4 * hid;
            Tensor gx = mk(dg);
            Dims dfull = din; dfull.c = dirs * hid;
            Tensor full = mk(dfull);
            // the tensor-core recurrences can emit the consumer's fp16 operand planes themselves (no k_split_f16 pass)
            const int ks_p = hid <= 32 ? 1 : hid <= 64 ? 2 : hid <= 128 ? 4 : 8;
            const bool generic_rec = w.whh_t && (hid > 256 || n.legacy == 2 || (getenv("KB_LSTM_GENERIC") && atoi(getenv("KB_LSTM_GENERIC")) != 0));   // writes fp32 only
            if (n.legacy == 2 && packed) throw Unsupported(n.name + ": 'PackedSequence' object has no attribute 'transpose' (the reference's ocropy cell cannot run on batches with seq_lens)");
            const bool tc_rec = (ks_p == 8 || ks_p == 1) && w.wpk && m->use_tc && !generic_rec && !(getenv("KB_LSTM_TC") && atoi(getenv("KB_LSTM_TC")) == 0);
            const bool lplanes = planes_hint && tc_rec && !n.summarize;
            planes_hint = false;
            if (lplanes) { full.hi = (__half *)ws->arena.alloc((size_t)full.numel() * 2); full.lo = (__half *)ws->arena.alloc((size_t)full.numel() * 2); }
            int *dl = packed ? dev_lens(lens) : nullptr;
            // per-step GEMM path (hidden > 256): gate pre-activations of one step, h and c state for both directions
            float *gen_g = nullptr, *gen_h = nullptr, *gen_c = nullptr;
            if (w.whh_t) {
                const size_t nsq = (size_t)(n.transpose ? x.n * x.w : x.n * x.h);
                gen_g = (float *)ws->arena.alloc(nsq * 4 * hid * sizeof(float));
                gen_h = (float *)ws->arena.alloc((size_t)dirs * nsq * hid * sizeof(float));
                gen_c = (float *)ws->arena.alloc((size_t)dirs * nsq * hid * sizeof(float));
            }
            if (full.numel()) { StageTimer tt(ws, st, n.name + ".xproj", !dry); gemm(x, w, nullptr, ACT_LINEAR, gx.p, x.h, x.w); }
            if (!dry && full.numel()) {
                StageTimer tt(ws, st, n.name + ".rec", true);
                LstmParams lp;
                lp.gx = gx.p; lp.whh = w.aux; lp.out = full.p; lp.lens = dl; lp.hid = hid; lp.dirs = dirs;
                if (!n.transpose) { lp.nseq = (int)(x.n * x.h); lp.T = (int)x.w; lp.q2 = 1; lp.s_outer = x.w; lp.s_inner = 0; lp.step = 1; }
                else { lp.nseq = (int)(x.n * x.w); lp.T = (int)x.h; lp.q2 = (int)x.w; lp.s_outer = x.h * x.w; lp.s_inner = 1; lp.step = x.w; }
                const int ks = hid <= 32 ? 1 : hid <= 64 ? 2 : hid <= 128 ? 4 : 8;
                lp.U = (hid + ks - 1) / ks;
                const bool generic = generic_rec;
                // tcgen05 recurrence (lstm_tc.cuh) for hidden sizes 129..256: 0.44 ms vs 0.69 ms for the CUDA-core kernel on cfg2 and
                // only 64 instead of 112 SMs; KB_LSTM_TC=0 selects the CUDA-core kernel (accurate expf/tanhf, fp32 FMA)
                const bool tc_on = w.wpk && m->use_tc && !(getenv("KB_LSTM_TC") && atoi(getenv("KB_LSTM_TC")) == 0);
                const bool rec_tc = ks == 8 && tc_on;
                if (generic) {
                    run_lstm_generic(n, w, lp, gen_g, gen_h, gen_c, lens);
                } else if (ks == 1 && tc_on) {
                    // hid <= 32 (blla's Lbx32 / Lby32): 64 sequences per CTA, W_hh in tensor memory, no cluster
                    ltc::LstmTcParams tp;
                    tp.gx = lp.gx; tp.wpk = (const uint16_t *)w.wpk; tp.out = (lplanes && !m->keep_fp32) ? nullptr : lp.out; tp.lens = lp.lens; tp.nseq = lp.nseq; tp.T = lp.T;
                    tp.out_hi = full.hi; tp.out_lo = full.lo;
                    tp.dbg = 0; tp.lpc = 0; tp.dbgbuf = nullptr; tp.hid = hid; tp.dirs = dirs; tp.U = hid; tp.q2 = lp.q2; tp.s_outer = lp.s_outer; tp.s_inner = lp.s_inner; tp.step = lp.step;
                    int snl = lp.nseq * dirs >= 64 * 2 * sm ? 64 : lp.nseq * dirs >= 16 * 4 * sm ? 32 : 16;      // fill the SMs first, then grow the CTAs
                    if (getenv("KB_LSTM_SNL")) snl = atoi(getenv("KB_LSTM_SNL"));
                    if (snl != 64 && snl != 32) snl = 16;
                    const dim3 grid((unsigned)((lp.nseq + snl - 1) / snl), (unsigned)dirs, 1);
                    if (getenv("KB_DEBUG")) fprintf(stderr, "[kb] %s: tcgen05 recurrence (single CTA, %d lines), %u CTAs, T=%d\n", n.name.c_str(), snl, grid.x * grid.y, lp.T);
                    if (snl == 64) ltc::k_lstm_rec_tc_small<64><<<grid, ltc::SmallCfg<64>::THREADS, ltc::SmallCfg<64>::SMEM_BYTES, st>>>(tp);
                    else if (snl == 32) ltc::k_lstm_rec_tc_small<32><<<grid, ltc::SmallCfg<32>::THREADS, ltc::SmallCfg<32>::SMEM_BYTES, st>>>(tp);
                    else ltc::k_lstm_rec_tc_small<16><<<grid, ltc::SmallCfg<16>::THREADS, ltc::SmallCfg<16>::SMEM_BYTES, st>>>(tp);
                    ++m->launches;
                    CK(cudaPeekAtLastError());
                } else if (rec_tc) {
                    ltc::LstmTcParams tp;
                    tp.gx = lp.gx; tp.wpk = (const uint16_t *)w.wpk; tp.out = (lplanes && !m->keep_fp32) ? nullptr : lp.out; tp.lens = lp.lens; tp.nseq = lp.nseq; tp.T = lp.T;
                    tp.out_hi = full.hi; tp.out_lo = full.lo;
                    tp.dbg = getenv("KB_LSTM_DBG") ? atoi(getenv("KB_LSTM_DBG")) : 0; tp.hid = hid; tp.dirs = dirs; tp.U = lp.U; tp.q2 = lp.q2; tp.s_outer = lp.s_outer; tp.s_inner = lp.s_inner; tp.step = lp.step;
                    // lines per cluster: 16 (two groups of 8); KB_LSTM_GL=16 selects 32 (two groups of 16: half the SMs per batch, but
                    // the longer epilogue stretches the per-step latency chain by 1.7x)
                    int gl = 8;                                  // measured on cfg2: 32 lines per cluster = 0.54 ms vs 0.31 ms, and no e2e gain
                    if (getenv("KB_LSTM_GL")) gl = atoi(getenv("KB_LSTM_GL")) == 16 ? 16 : 8;
                    // Groups of 8 lines per cluster.  Two (16 lines on 8 SMs, the groups alternating on the tensor pipe) give the shortest
                    // recurrence (cfg2: 0.311 ms) and are what a synchronous call gets.  Three free-running groups (24 lines on 8 SMs) take
                    // 0.356 ms but a third fewer SMs, which is worth +6.5 % lines/s when several batches are in flight (bench.py, r02):
                    // the asynchronous slots use that.  Four groups (one epilogue warp per TMEM lane quarter) take 0.52 ms: no gain, kept as a switch.
                    // KB_LSTM_NG=2|3|4 overrides.  DESIGN.md 4.2.
                    int ng = (gl == 8 && ws->index > 0 && (lp.nseq + 23) / 24 < (lp.nseq + 15) / 16) ? 3 : 2;      // only where it saves clusters
                    if (gl == 8 && getenv("KB_LSTM_NG")) { const int e = atoi(getenv("KB_LSTM_NG")); ng = e == 3 || e == 4 ? e : 2; }     // 4: one epilogue warp per TMEM lane quarter and group
                    tp.alt = getenv("KB_LSTM_ALT") ? atoi(getenv("KB_LSTM_ALT")) : (ng == 2);
                    const int nl = ng * gl;
                    // Lines per cluster.  Measured on cfg2 (tools/rec_ab.py): 16, 12 and 10 lines per cluster all take 0.30 ms - the time step
                    // is a latency chain (~500 MMAs + ~800 epilogue + ~1100-1300 remote delivery), not a function of the bytes sent - so a
                    // synchronous call gains nothing from spreading over more clusters and every call keeps full clusters (fewest SMs per
                    // batch).  KB_LSTM_LPC overrides (ragged last cluster / tests).
                    int lpc = (lp.nseq + ((lp.nseq + nl - 1) / nl) - 1) / std::max(1, (lp.nseq + nl - 1) / nl);      // the same number of clusters, evenly filled
                    lpc = std::min(nl, std::max(1, lpc));
                    if (gl == 8) {
                        if (m->max_clusters_tc < 0) {
                            cudaLaunchConfig_t q = {};
                            q.gridDim = dim3(ltc::LCS * 32, 1, 1); q.blockDim = dim3(ltc::LTHREADS, 1, 1); q.dynamicSmemBytes = ltc::ClusterCfg<8, 2>::SMEM_BYTES;
                            cudaLaunchAttribute qa[1]; qa[0].id = cudaLaunchAttributeClusterDimension;
                            qa[0].val.clusterDim.x = ltc::LCS; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
                            q.attrs = qa; q.numAttrs = 1;
                            int ncl = 0;
                            if (cudaOccupancyMaxActiveClusters(&ncl, ltc::k_lstm_rec_tc<8, 2>, &q) != cudaSuccess) { cudaGetLastError(); ncl = 8; }
                            m->max_clusters_tc = std::max(ncl, 1);
                        }
                        if (getenv("KB_LSTM_LPC")) lpc = std::min(nl, std::max(1, atoi(getenv("KB_LSTM_LPC"))));
                    }
                    tp.lpc = lpc;
                    cudaLaunchConfig_t tcfg = {};
                    tcfg.gridDim = dim3((unsigned)(ltc::LCS * ((lp.nseq + lpc - 1) / lpc)), (unsigned)dirs, 1);
                    tcfg.blockDim = dim3(ltc::lthreads(ng, ng == 4 ? 1 : 2), 1, 1);
                    tcfg.dynamicSmemBytes = gl == 16 ? ltc::ClusterCfg<16, 2>::SMEM_BYTES : (ng == 4 ? ltc::ClusterCfg<8, 4, 1>::SMEM_BYTES : ng == 3 ? ltc::ClusterCfg<8, 3>::SMEM_BYTES : ltc::ClusterCfg<8, 2>::SMEM_BYTES); tcfg.stream = st;
                    cudaLaunchAttribute tat[1];
                    tat[0].id = cudaLaunchAttributeClusterDimension;
                    tat[0].val.clusterDim.x = ltc::LCS; tat[0].val.clusterDim.y = 1; tat[0].val.clusterDim.z = 1;
                    tcfg.attrs = tat; tcfg.numAttrs = 1;
                    if (getenv("KB_DEBUG")) fprintf(stderr, "[kb] %s: tcgen05 recurrence, %u clusters of 8 CTAs x %d lines (max co-resident %d), T=%d\n", n.name.c_str(), tcfg.gridDim.x / 8 * dirs, lpc, m->max_clusters_tc, lp.T);
                    tp.dbgbuf = nullptr;
                    if (tp.dbg & 1) { CK(cudaMalloc((void **)&tp.dbgbuf, 96 * sizeof(long long))); CK(cudaMemset(tp.dbgbuf, 0, 96 * sizeof(long long))); }
                    if (gl == 16) CK(cudaLaunchKernelEx(&tcfg, ltc::k_lstm_rec_tc<16, 2>, tp));
                    else if (ng == 4) CK(cudaLaunchKernelEx(&tcfg, ltc::k_lstm_rec_tc<8, 4, 1>, tp));
                    else if (ng == 3) CK(cudaLaunchKernelEx(&tcfg, ltc::k_lstm_rec_tc<8, 3>, tp));
                    else CK(cudaLaunchKernelEx(&tcfg, ltc::k_lstm_rec_tc<8, 2>, tp));
                    if (tp.dbg & 1) {
                        long long hb[96];
                        CK(cudaMemcpy(hb, tp.dbgbuf, sizeof(hb), cudaMemcpyDeviceToHost));
                        cudaFree(tp.dbgbuf);
                        const long long t0 = hb[1];
                        for (int i = 0; i < 8; ++i) {
                            const long long *d = hb + i * 12;
                            fprintf(stderr, "[tcrec] step %d group %d: issuer: wait_from %lld  h_ready %lld  issued %lld | epilogue: loop_top %lld  mma_done %lld  tmem_read %lld  act_end %lld  cell+gather %lld  sent %lld  loop_end %lld\n",
                                    100 + i / 2, i & 1, d[0] - t0, d[1] - t0, d[2] - t0, d[3] - t0, d[4] - t0, d[5] - t0, d[6] - t0, d[7] - t0, d[8] - t0, d[9] - t0);
                        }
                    }
                    ++m->launches;
                    CK(cudaPeekAtLastError());
                } else {
                const int BL = 64 / ks;
                const int nchunks = (lp.nseq + BL - 1) / BL;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)(ks * nchunks), (unsigned)dirs, 1);
                cfg.blockDim = dim3(256, 1, 1);
                cfg.dynamicSmemBytes = 0; cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = (unsigned)ks; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                bool lines10 = false;
                if (ks == 8) {
                    // only ~15 clusters of 8 CTAs are co-resident on a B200; with 10 lines per cluster cfg2's 64 lines x 2
                    // directions need 14 clusters (one wave) instead of 16 (two waves)
                    if (m->max_clusters8 < 0) {
                        int ncl = 0;
                        if (cudaOccupancyMaxActiveClusters(&ncl, k_lstm_rec<8, 8>, &cfg) != cudaSuccess) { cudaGetLastError(); ncl = 0; }
                        m->max_clusters8 = ncl;
                    }
                    const int c8 = nchunks * dirs, c10 = ((lp.nseq + 9) / 10) * dirs;
                    const int cap = std::max(m->max_clusters8, 1);
                    lines10 = (c10 + cap - 1) / cap < (c8 + cap - 1) / cap;
                    if (getenv("KB_LSTM_LINES")) lines10 = atoi(getenv("KB_LSTM_LINES")) == 10;
                    if (lines10) cfg.gridDim = dim3((unsigned)(ks * ((lp.nseq + 9) / 10)), (unsigned)dirs, 1);
                    if (getenv("KB_DEBUG"))
                        fprintf(stderr, "[kb] %s: lstm rec ks=8 lines/cluster=%d clusters=%u, T=%d, max co-resident clusters=%d\n", n.name.c_str(),
                                lines10 ? 10 : 8, cfg.gridDim.x / 8 * dirs, lp.T, m->max_clusters8);
                }
                switch (ks) {
                case 1: CK(cudaLaunchKernelEx(&cfg, k_lstm_rec<1, 8>, lp)); break;
                case 2: CK(cudaLaunchKernelEx(&cfg, k_lstm_rec<2, 8>, lp)); break;
                case 4: CK(cudaLaunchKernelEx(&cfg, k_lstm_rec<4, 8>, lp)); break;
                default:
                    if (lines10) CK(cudaLaunchKernelEx(&cfg, k_lstm_rec<8, 10>, lp));
                    else CK(cudaLaunchKernelEx(&cfg, k_lstm_rec<8, 8>, lp));
                    break;
                }
                ++m->launches;
                CK(cudaPeekAtLastError());
                }
            }
            if (n.summarize) {
                y = mk(dout);
                if (!dry && y.numel()) LAUNCH(m, k_take_last, grid1d(y.numel(), 256, sm), 256, 0, st, full.p, y.p, (int)full.n, (int)full.h, (int)full.w, (int)full.c, n.transpose ? 1 : 0);
            } else y = full;
            break;
        }
        default: throw Unsupported("layer kind not implemented");
        }
        if (lens.has) for (auto &l : lens.v) l = leaf_len(n, l, din, dout);
        return y;
    }

    static bool passthrough(const Node &c) { return c.kind == K_DROPOUT || c.kind == K_IDENTITY; }
    static bool pool22(const Node &c) { return c.kind == K_POOL && c.kh == 2 && c.kw == 2 && c.sy == 2 && c.sx == 2; }
    // next layer after position j that is not an eval-mode no-op
    static const Node *next_real(const Node &series, size_t j, size_t *pos = nullptr) {
        while (j < series.children.size() && passthrough(*series.children[j])) ++j;
        if (pos) *pos = j;
        return j < series.children.size() ? series.children[j].get() : nullptr;
    }
    void advance_lens(const Node &c, Lens &lens, const Dims &din, const Dims &dout) {
        if (lens.has) for (auto &l : lens.v) l = leaf_len(c, l, din, dout);
    }

    // does the layer at/after position j of `series` read fp16 operand planes of a tensor with dims d?  (tensor-core conv, or an
    // LSTM / Linear whose projection runs on k_gemm_tc)
    bool wants_planes(const Node &series, size_t j, const Dims &d) const {
        const Node *nx = next_real(series, j);
        if (!nx || !m->use_tc) return false;
        if (nx->kind == K_CONV) return (m->fuse_mask & 2) && tc_conv_eligible(*nx, d);
        if ((nx->kind == K_LSTM && !nx->legacy) || nx->kind == K_LINEAR) {
            const LeafWeights &nw = m->lw[nx->leaf_index];
            const long long M = (long long)d.n * d.h * d.w;
            return nw.b_hi && (d.c % 8) == 0 && d.c >= 32 && nw.ncols >= 64 && M >= 128 && !(nx->kind == K_LINEAR && nx->aug);
        }
        return false;
    }

    // Fused layer groups.  Returns the number of children of `series` consumed (0 = no pattern applies); on success `cur`
    // and `lens` are advanced past the group.  Layers inside a group do not materialise (no kb_debug_layer_output tap).
    size_t try_fuse(const Node &series, size_t i, Tensor &cur, Lens &lens) {
        const Node &c0 = *series.children[i];
        // ---- P1: conv(Cin = 1, stride 1, dilation 1) -> [Dropout/Identity]* -> MaxPool 2x2/2  (stencil + pool in one pass)
        if ((m->fuse_mask & 1) && c0.kind == K_CONV && c0.cin == 1 && cur.c == 1 && c0.sy == 1 && c0.sx == 1 && c0.dy == 1 && c0.dx == 1 &&
            (c0.cout == 8 || c0.cout == 16 || c0.cout == 32 || c0.cout == 64) &&
            (c0.act == ACT_RELU || c0.act == ACT_LINEAR || c0.act == ACT_SIGMOID_LOGITS)) {
            size_t j; const Node *pl = next_real(series, i + 1, &j);
            if (!pl || !pool22(*pl)) return 0;
            const Dims din = dims_of(cur);
            const Dims dconv = leaf_dims(c0, din);
            if (dconv.h < 2 || dconv.w < 2) return 0;
            const Dims dpool = leaf_dims(*pl, dconv);
            const LeafWeights &w = m->lw[c0.leaf_index];
            Tensor y = mk(dpool);
            // a tensor-core conv right behind wants the operand planes
            const Node *nx = next_real(series, j + 1);
            const bool planes = m->use_tc && (m->fuse_mask & 2) && nx && nx->kind == K_CONV && tc_conv_eligible(*nx, dpool);
            if (planes) { y.hi = (__half *)ws->arena.alloc((size_t)y.numel() * 2); y.lo = (__half *)ws->arena.alloc((size_t)y.numel() * 2); }
            if (!dry && y.numel()) {
                StageTimer tt(ws, st, c0.name + "+" + pl->name, true);
                Conv1PoolParams cp;
                cp.x = cur.p; cp.wt = w.wt; cp.bias = w.bias; cp.y = (planes && !m->keep_fp32) ? nullptr : y.p; cp.y_hi = y.hi; cp.y_lo = y.lo; cp.flag = ws->d_flag;
                cp.N = (int)cur.n; cp.H = (int)cur.h; cp.W = (int)cur.w; cp.Cout = c0.cout; cp.Ncp = w.ncp; cp.kh = c0.kh; cp.kw = c0.kw;
                cp.py = c0.py; cp.px = c0.px; cp.Hp = (int)dpool.h; cp.Wp = (int)dpool.w; cp.act = c0.act;
                const int cgroups = c0.cout / 8, ppb = 256 / cgroups;
                const int tw = 2 * ppb + c0.kw - 1, th = c0.kh + 1;
                const size_t smem = ((size_t)((th * tw + 3) & ~3) + (size_t)c0.kh * c0.kw * c0.cout) * sizeof(float);
                if (smem > 48 * 1024) throw Unsupported(c0.name + ": filter bank too large for the fused stencil kernel");
                const char *c1env = getenv("KB_CONV1");
                if (w.c1_pk && m->use_tc && !m->force_ffma && !(c1env && (strcmp(c1env, "generic") == 0 || strcmp(c1env, "ffma") == 0))) {
                    // tcgen05 version: the nine taps are one K16 step (conv1_tc.cuh).  Tile shape with the smaller padding waste.
                    c1tc::Conv1TcParams tp;
                    tp.x = cur.p; tp.wpk = w.c1_pk; tp.bias = w.bias; tp.y = cp.y; tp.y_hi = cp.y_hi; tp.y_lo = cp.y_lo; tp.flag = ws->d_flag;
                    tp.N = cp.N; tp.H = cp.H; tp.W = cp.W; tp.Cout = cp.Cout; tp.Hp = cp.Hp; tp.Wp = cp.Wp; tp.act = cp.act;
                    const long long w128 = (long long)((cp.Wp + 127) / 128) * 128 * cp.Hp, w64 = (long long)((cp.Wp + 63) / 64) * 64 * ((cp.Hp + 1) / 2) * 2;
                    if (w64 < w128) { tp.tpw = 64; tp.tph = 2; } else { tp.tpw = 128; tp.tph = 1; }
                    tp.tiles_w = (cp.Wp + tp.tpw - 1) / tp.tpw; tp.tiles_h = (cp.Hp + tp.tph - 1) / tp.tph;
                    const long long ntiles = (long long)tp.N * tp.tiles_w * tp.tiles_h;
                    if (cp.Cout == 32) LAUNCH(m, c1tc::k_conv1_tc<32>, (unsigned)std::min<long long>(ntiles, 2LL * m->sm_count), c1tc::C1_THREADS, c1tc::conv1_tc_smem(32), st, tp);
                    else LAUNCH(m, c1tc::k_conv1_tc<16>, (unsigned)std::min<long long>(ntiles, 2LL * m->sm_count), c1tc::C1_THREADS, c1tc::conv1_tc_smem(16), st, tp);
                } else if (c0.kh == 3 && c0.kw == 3 && (w.ncp % 4) == 0 && !(c1env && strcmp(c1env, "generic") == 0)) {
                    // register-resident 3x3 filter bank, strips of 8 pooled rows per block
                    constexpr int RP = 8;
                    dim3 grid3((unsigned)((dpool.w + ppb - 1) / ppb), (unsigned)((dpool.h + RP - 1) / RP), (unsigned)dpool.n);
                    const size_t sm3 = (size_t)2 * 4 * (2 * ppb + 2) * sizeof(float);
                    switch (c0.cout) {                                         // patch elements per thread = ceil(4 * (2 * ppb + 2) / 256)
                    case 64: LAUNCH(m, (k_conv1_pool33<RP, 2>), grid3, 256, sm3, st, cp); break;
                    case 32: LAUNCH(m, (k_conv1_pool33<RP, 3>), grid3, 256, sm3, st, cp); break;
                    case 16: LAUNCH(m, (k_conv1_pool33<RP, 5>), grid3, 256, sm3, st, cp); break;
                    default: LAUNCH(m, (k_conv1_pool33<RP, 9>), grid3, 256, sm3, st, cp); break;
                    }
                } else {
                    dim3 grid((unsigned)((dpool.w + ppb - 1) / ppb), (unsigned)dpool.h, (unsigned)dpool.n);
                    LAUNCH(m, k_conv1_pool, grid, 256, smem, st, cp);
                }
            }
            advance_lens(c0, lens, din, dconv);
            advance_lens(*pl, lens, dconv, dpool);
            cur = y;
            if (!dry) ws->taps[pl->name] = y;
            return j + 1 - i;
        }
        if ((m->fuse_mask & 4) && c0.kind == K_CONV && m->lw[c0.leaf_index].s2d) { const size_t u = fuse_conv_s2d(series, i, cur, lens); if (u) return u; }
        if ((m->fuse_mask & 2) && c0.kind == K_CONV) return fuse_conv_tc(series, i, cur, lens);
        return 0;
    }
    // tensor-core convolution (conv_tc.cuh): stride-1, undilated, Cin a multiple of 32, Cout a multiple of 16 up to 256
    bool tc_conv_eligible(const Node &c, const Dims &in) const {
        if (!m->use_tc || c.kind != K_CONV || in.c != c.cin || (in.c % 32) != 0) return false;
        const LeafWeights &w = m->lw[c.leaf_index];
        if (!w.c_hi || w.s2d || c.sy != 1 || c.sx != 1) return false;
        if (!(c.act == ACT_RELU || c.act == ACT_LINEAR || c.act == ACT_SIGMOID_LOGITS || c.act == ACT_TANH || c.act == ACT_LEAKY)) return false;
        if (in.h < 1 || in.w < 1) return false;
        return ctc::conv_tc_plan(c.kh, c.kw, c.cout, nullptr, nullptr, nullptr) <= 227 * 1024;
    }
    // stride-2 convolution as a stride-1 convolution over the space-to-depth planes (weights repacked at finalize)
    bool s2d_conv_eligible(const Node &c, const Dims &in) const {
        if (!m->use_tc || !m->fuse || !(m->fuse_mask & 4) || c.kind != K_CONV || in.c != c.cin) return false;
        const LeafWeights &w = m->lw[c.leaf_index];
        if (!w.c_hi || !w.s2d) return false;
        if (!(c.act == ACT_RELU || c.act == ACT_LINEAR || c.act == ACT_SIGMOID_LOGITS || c.act == ACT_TANH || c.act == ACT_LEAKY)) return false;
        return in.h >= 1 && in.w >= 1;
    }
    bool first_is_s2d(const Node &root, const Dims &in) const {
        return root.kind == K_SERIES && !root.children.empty() && s2d_conv_eligible(*root.children[0], in);
    }
    static bool fold_h(const Node &c) { return c.kind == K_RESHAPE && c.rs_src == 2 && c.rs_a == 1 && c.rs_b == -1 && c.rs_high == 2 && c.rs_low == 1; }

    // ---- P3: conv(stride 2) on tcgen05 via space-to-depth: the planes come from the GroupNorm in front (k_gn_apply4), from the
    // NCHW network input, or from k_s2d_planes over an NHWC activation
    size_t fuse_conv_s2d(const Node &series, size_t i, Tensor &cur, Lens &lens) {
        const Node &c0 = *series.children[i];
        const Dims din = dims_of(cur);
        if (!s2d_conv_eligible(c0, din)) return 0;
        const LeafWeights &w = m->lw[c0.leaf_index];
        const Dims dconv = leaf_dims(c0, din);
        const int64_t Hb = (din.h + 1) / 2, Wb = (din.w + 1) / 2;
        const int cs = w.s_cs;
        Tensor y = mk(dconv);
        const bool planes = wants_planes(series, i + 1, dconv);
        if (planes) { y.hi = (__half *)ws->arena.alloc((size_t)y.numel() * 2); y.lo = (__half *)ws->arena.alloc((size_t)y.numel() * 2); }
        __half *x_hi = cur.s_hi, *x_lo = cur.s_lo;
        const bool have = x_hi && cur.s2d_c == cs;
        if (!have) {
            const size_t se = (size_t)din.n * Hb * Wb * cs;
            x_hi = (__half *)ws->arena.alloc(se * 2); x_lo = (__half *)ws->arena.alloc(se * 2);
        }
        if (!dry && y.numel()) {
            StageTimer tt(ws, st, c0.name, true);
            if (!have) {
                const float *src = cur.nchw ? cur.nchw : cur.p;
                LAUNCH(m, k_s2d_planes, grid1d((long long)din.n * Hb * Wb * (cs / 8), 256, m->sm_count), 256, 0, st, src, cur.nchw ? 1 : 0, x_hi, x_lo,
                       (int)din.n, (int)din.c, (int)din.h, (int)din.w, (int)Hb, (int)Wb, cs, ws->d_flag);
            }
            ctc::ConvTcParams cp;
            cp.bias = w.bias; cp.y = (planes && !m->keep_fp32) ? nullptr : y.p; cp.y_hi = y.hi; cp.y_lo = y.lo; cp.flag = ws->d_flag;
            cp.N = (int)din.n; cp.Ho = (int)dconv.h; cp.Wo = (int)dconv.w; cp.Cout = c0.cout; cp.kh = w.s_th; cp.kw = w.s_tw; cp.py = w.s_py; cp.px = w.s_px;
            cp.act = c0.act; cp.pool = 0;
            cp.out_h = (int)dconv.h; cp.out_w = (int)dconv.w;
            cp.items_h = (int)((dconv.h + 1) / 2);
            cp.items_w = (int)((dconv.w + ctc::TW - 1) / ctc::TW);
            cp.sN = dconv.h * dconv.w * dconv.c; cp.sH = dconv.w * dconv.c; cp.sW = dconv.c;
            const bool wres_ok = !(getenv("KB_CONV_WRES") && atoi(getenv("KB_CONV_WRES")) == 0);
            const size_t smem = ctc::conv_tc_plan(w.s_th, w.s_tw, c0.cout, &cp.CT, &cp.nstb, &cp.a_row_bytes, &cp.a_sets, wres_ok ? cs / 32 : 0, &cp.w_res);
            cp.NC = cs / 32; cp.items_c = c0.cout / cp.CT;
            cp.acc_sets = 8 * cp.CT <= 512 ? 2 : 1;
            CUtensorMap tx_hi, tx_lo, tw_hi, tw_lo;
            const uint32_t box_w = (uint32_t)(ctc::TW + w.s_tw - 1);
            if (!ctc::make_map_nhwc(&tx_hi, x_hi, (uint64_t)din.n, (uint64_t)Hb, (uint64_t)Wb, (uint64_t)cs, box_w) ||
                !ctc::make_map_nhwc(&tx_lo, x_lo, (uint64_t)din.n, (uint64_t)Hb, (uint64_t)Wb, (uint64_t)cs, box_w) ||
                !tc::make_map_2d(&tw_hi, w.c_hi, (uint64_t)w.s_th * w.s_tw * cp.NC * c0.cout, 32, (uint32_t)cp.CT) ||
                !tc::make_map_2d(&tw_lo, w.c_lo, (uint64_t)w.s_th * w.s_tw * cp.NC * c0.cout, 32, (uint32_t)cp.CT))
                throw CudaError("cuTensorMapEncodeTiled failed (strided conv)");
            const int nitems = cp.N * cp.items_h * cp.items_w * cp.items_c;
            if (nitems > 0) LAUNCH(m, ctc::k_conv_tc, (unsigned)std::min(nitems, m->sm_count), ctc::CTHREADS, smem, st, tx_hi, tx_lo, tw_hi, tw_lo, cp);
        }
        advance_lens(c0, lens, din, dconv);
        cur = y;
        if (!dry) ws->taps[c0.name] = y;
        return 1;
    }

    // ---- P2: conv(Cin = 32, stride 1) -> [Do]* -> [MaxPool 2x2/2] -> [Do]* -> [S fold of H into features] on tcgen05
    size_t fuse_conv_tc(const Node &series, size_t i, Tensor &cur, Lens &lens) {
        const Node &c0 = *series.children[i];
        const Dims din = dims_of(cur);
        if (!tc_conv_eligible(c0, din)) return 0;
        const Dims dconv = leaf_dims(c0, din);
        size_t j = i + 1, jpool = 0, jfold = 0;
        const Node *pl = nullptr, *fd = nullptr;
        { size_t k; const Node *nx = next_real(series, j, &k); if (nx && pool22(*nx) && dconv.h >= 2 && dconv.w >= 2) { pl = nx; jpool = k; j = k + 1; } }
        Dims dpost = pl ? leaf_dims(*pl, dconv) : dconv;
        { size_t k; const Node *nx = next_real(series, j, &k); if (nx && fold_h(*nx)) { fd = nx; jfold = k; j = k + 1; } }
        Dims dout = fd ? leaf_dims(*fd, dpost) : dpost;
        (void)jpool; (void)jfold;
        Tensor y = mk(dout);
        // consumer wants operand planes?  (another tensor-core conv, or an LSTM / Linear whose projection runs on k_gemm_tc)
        const bool planes = wants_planes(series, j, dout);
        if (planes) { y.hi = (__half *)ws->arena.alloc((size_t)y.numel() * 2); y.lo = (__half *)ws->arena.alloc((size_t)y.numel() * 2); }
        __half *x_hi = cur.hi, *x_lo = cur.lo;
        if (!x_hi) { x_hi = (__half *)ws->arena.alloc((size_t)cur.numel() * 2); x_lo = (__half *)ws->arena.alloc((size_t)cur.numel() * 2); }
        if (!dry && y.numel()) {
            std::string nm = c0.name; if (pl) nm += "+" + pl->name; if (fd) nm += "+" + fd->name;
            StageTimer tt(ws, st, nm, true);
            if (!cur.hi) LAUNCH(m, tc::k_split_f16, grid1d(cur.numel() / 4, 256, m->sm_count), 256, 0, st, cur.p, x_hi, x_lo, (long long)(cur.numel() / 4), ws->d_flag);
            const LeafWeights &w = m->lw[c0.leaf_index];
            ctc::ConvTcParams cp;
            cp.bias = w.bias; cp.y = (planes && !m->keep_fp32) ? nullptr : y.p; cp.y_hi = y.hi; cp.y_lo = y.lo; cp.flag = ws->d_flag;
            cp.N = (int)cur.n; cp.Ho = (int)dconv.h; cp.Wo = (int)dconv.w; cp.Cout = c0.cout; cp.kh = c0.kh; cp.kw = c0.kw; cp.py = c0.py; cp.px = c0.px;
            cp.act = c0.act; cp.pool = pl ? 1 : 0;
            cp.out_h = (int)dpost.h; cp.out_w = (int)dpost.w;
            cp.items_h = pl ? (int)dpost.h : (int)((dconv.h + 1) / 2);
            cp.items_w = (int)((dconv.w + ctc::TW - 1) / ctc::TW);
            if (fd) { cp.sN = dpost.w * dpost.h * dpost.c; cp.sW = dpost.h * dpost.c; cp.sH = dpost.c; }
            else { cp.sN = dpost.h * dpost.w * dpost.c; cp.sH = dpost.w * dpost.c; cp.sW = dpost.c; }
            const bool wres_ok = !(getenv("KB_CONV_WRES") && atoi(getenv("KB_CONV_WRES")) == 0);
            const size_t smem = ctc::conv_tc_plan(c0.kh, c0.kw, c0.cout, &cp.CT, &cp.nstb, &cp.a_row_bytes, &cp.a_sets, wres_ok ? c0.cin / 32 : 0, &cp.w_res);
            cp.NC = c0.cin / 32; cp.items_c = c0.cout / cp.CT;
            cp.acc_sets = 8 * cp.CT <= 512 ? 2 : 1;
            CUtensorMap tx_hi, tx_lo, tw_hi, tw_lo;
            const uint32_t box_w = (uint32_t)(ctc::TW + c0.kw - 1);
            if (!ctc::make_map_nhwc(&tx_hi, x_hi, (uint64_t)cur.n, (uint64_t)cur.h, (uint64_t)cur.w, (uint64_t)cur.c, box_w) ||
                !ctc::make_map_nhwc(&tx_lo, x_lo, (uint64_t)cur.n, (uint64_t)cur.h, (uint64_t)cur.w, (uint64_t)cur.c, box_w) ||
                !tc::make_map_2d(&tw_hi, w.c_hi, (uint64_t)c0.kh * c0.kw * cp.NC * c0.cout, 32, (uint32_t)cp.CT) ||
                !tc::make_map_2d(&tw_lo, w.c_lo, (uint64_t)c0.kh * c0.kw * cp.NC * c0.cout, 32, (uint32_t)cp.CT))
                throw CudaError("cuTensorMapEncodeTiled failed (conv)");
            const int nitems = cp.N * cp.items_h * cp.items_w * cp.items_c;
            if (nitems > 0) LAUNCH(m, ctc::k_conv_tc, (unsigned)std::min(nitems, m->sm_count), ctc::CTHREADS, smem, st, tx_hi, tx_lo, tw_hi, tw_lo, cp);
        }
        advance_lens(c0, lens, din, dconv);
        if (pl) advance_lens(*pl, lens, dconv, dpost);
        if (fd) advance_lens(*fd, lens, dpost, dout);
        cur = y;
        if (!dry) ws->taps[fd ? fd->name : (pl ? pl->name : c0.name)] = y;
        return j - i;
    }

    Tensor run(const Node &n, const Tensor &x, Lens &lens) {
        if (n.kind == K_SERIES) {
            Tensor cur = x;
            for (size_t i = 0; i < n.children.size();) {
                size_t used = 0;
                if (m->fuse) used = try_fuse(n, i, cur, lens);
                if (used) { i += used; continue; }
                if (m->fuse && (n.children[i]->kind == K_GN || n.children[i]->kind == K_LSTM)) {
                    Dims dn = dims_of(cur);
                    if (n.children[i]->kind == K_LSTM) { Lens ltmp; dn = leaf_dims(*n.children[i], dn); (void)ltmp; }
                    planes_hint = wants_planes(n, i + 1, dn);
                    if (n.children[i]->kind == K_GN) { const Node *nx = next_real(n, i + 1); s2d_hint = nx && s2d_conv_eligible(*nx, dn) && m->lw[nx->leaf_index].s_cs == 4 * dn.c; }
                }
                cur = run(*n.children[i], cur, lens);
                planes_hint = false; s2d_hint = false;
                ++i;
            }
            return cur;
        }
        if (n.kind == K_PARALLEL) {
            std::vector<Tensor> outs; Lens last = lens; int64_t ctot = 0;
            for (auto &c : n.children) {
                Lens l = lens;
                outs.push_back(run(*c, x, l));
                last = l; ctot += outs.back().c;
            }
            for (auto &o : outs)
                if (o.n != outs[0].n || o.h != outs[0].h || o.w != outs[0].w) throw ShapeError("Output shape in parallel block not equal!");
            Dims d = dims_of(outs[0]); d.c = ctot;
            Tensor y = mk(d);
            int64_t off = 0;
            for (auto &o : outs) {
                if (!dry && o.numel()) LAUNCH(m, k_concat, grid1d(o.numel(), 256, m->sm_count), 256, 0, st, o.p, y.p, (long long)(o.n * o.h * o.w), (int)o.c, (int)ctot, (int)off);
                off += o.c;
            }
            lens = last;
            return y;
        }
        // leaf
        Tensor y;
        {
            StageTimer tt(ws, st, n.name, !dry && n.kind != K_LSTM && n.kind != K_DROPOUT && n.kind != K_IDENTITY);
            y = leaf(n, x, lens);
        }
        if (!dry) ws->taps[n.name] = y;
        return y;
    }
};

// Every ABI entry point that selects the model's device restores the caller's current device on the way out: torch (and any
// other runtime-API user in the process) reads the current device of the calling thread.
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static void ensure_ready(kb_model *m) {
    if (!m->finalized) throw SpecError("model not finalized: call kb_model_finalize() after loading weights");
    CK(cudaSetDevice(m->device));
}

// device-side pieces of a workspace that every call needs: range flag pair (and, for async slots, stream + events)
static void ensure_workspace(kb_model *m, Workspace *ws) {
    (void)m;
    if (!ws->d_flag) { CK(cudaMalloc((void **)&ws->d_flag, sizeof(int))); CK(cudaMemset(ws->d_flag, 0, sizeof(int))); }
    if (!ws->h_flag) { CK(cudaHostAlloc((void **)&ws->h_flag, sizeof(int), cudaHostAllocDefault)); *ws->h_flag = 0; }
    if (ws->index > 0 && !ws->stream) {
        CK(cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ws->done, cudaEventBlockingSync | cudaEventDisableTiming));      // kb_wait sleeps instead of spinning
        CK(cudaEventCreateWithFlags(&ws->input_ready, cudaEventDisableTiming));
    }
}

struct ForwardResult { Tensor y; Lens lens; };

// the network's last layer if it is a Linear sitting directly at the end of the top-level series (the recognition head)
static const Node *final_linear(const Node &root) {
    if (root.kind != K_SERIES || root.children.empty()) return nullptr;
    const Node *last = root.children.back().get();
    return last->kind == K_LINEAR ? last : nullptr;
}

// Stages the NCHW input into the arena as NHWC, runs the net.  `extra_bytes`: additional arena space the
// caller will allocate after the forward (decode buffers etc.).
static ForwardResult forward_impl(kb_model *m, Workspace *ws, const float *x, int x_on_device, int n, int h, int w, const int32_t *widths,
                                  cudaStream_t st, size_t extra_bytes, Exec::ArgmaxOut *amx = nullptr) {
    const Plan &pl = *m->plan;
    const int C = pl.input[1];
    { const char *e = getenv("KB_GEMM"); m->use_tc = !(e && strcmp(e, "ffma") == 0) && !m->force_ffma; }
    { const char *e = getenv("KB_KEEP_FP32"); m->keep_fp32 = e && strcmp(e, "0") != 0; }
    { const char *e = getenv("KB_FUSE"); m->fuse_mask = e ? atoi(e) : 7; m->fuse = m->fuse_mask != 0; }
    if (n <= 0 || h <= 0 || w <= 0) throw ShapeError("empty input batch");
    if (pl.input[2] > 0 && h != pl.input[2] && pl.input[2] != 1)
        ;   // the reference does not check the declared height either; convs accept any H
    Lens lens0;
    if (widths) { lens0.has = true; lens0.v.assign(widths, widths + n); }
    const size_t in_elems = (size_t)n * C * h * w;
    // ---- pass 1: plan the arena
    const auto prof_t0 = std::chrono::steady_clock::now();
    size_t need;
    bool first_s2d = false;
    {
        Arena saved = ws->arena;
        ws->arena.dry = true; ws->arena.off = 0;
        Exec ex{m, ws, st, true};
        ex.amx = amx; ex.final_leaf = final_linear(*pl.root);
        Dims d0; d0.n = n; d0.c = C; d0.h = h; d0.w = w;
        first_s2d = C > 1 && ex.first_is_s2d(*pl.root, d0);     // the first layer reads the NCHW input itself (space-to-depth planes)
        if (!x_on_device || C > 1) ws->arena.alloc(in_elems * sizeof(float));      // staging of the raw input
        if (C > 1 && !first_s2d) ws->arena.alloc(in_elems * sizeof(float));
        Tensor t; t.p = nullptr; t.n = n; t.c = C; t.h = h; t.w = w;
        Lens l = lens0;
        try { ex.run(*pl.root, t, l); } catch (...) { ws->arena = saved; throw; }
        need = ws->arena.off + extra_bytes + (1 << 20);
        ws->arena = saved;
    }
    if (need > ws->arena.cap) {
        CK(cudaStreamSynchronize(st));
        if (ws->arena.base) CK(cudaFree(ws->arena.base));
        ws->arena.base = nullptr; ws->arena.cap = 0;
        size_t cap = need + need / 4;
        CK(cudaMalloc((void **)&ws->arena.base, cap));
        ws->arena.cap = cap;
    }
    ws->arena.dry = false; ws->arena.off = 0;
    ws->taps.clear();
    ws->n_stages = 0;
    const auto prof_t1 = std::chrono::steady_clock::now();
    m->prof_us[0] += std::chrono::duration<double, std::micro>(prof_t1 - prof_t0).count();
    // ---- stage input
    Tensor t; t.n = n; t.c = C; t.h = h; t.w = w;
    const float *src = x;
    std::unique_ptr<StageTimer> t_in(new StageTimer(ws, st, "stage_in", true));
    if (!x_on_device) {
        float *stg = (float *)ws->arena.alloc(in_elems * sizeof(float));
        CK(cudaMemcpyAsync(stg, x, in_elems * sizeof(float), cudaMemcpyHostToDevice, st));
        src = stg;
    } else if (C > 1) ws->arena.alloc(in_elems * sizeof(float));   // keep offsets identical to the dry pass
    if (C > 1 && first_s2d) { t.p = nullptr; t.nchw = src; }
    else if (C > 1) {
        float *nhwc = (float *)ws->arena.alloc(in_elems * sizeof(float));
        dim3 grid((unsigned)((h * (long long)w + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)n);
        LAUNCH(m, k_transpose, grid, dim3(32, 8), 0, st, src, nhwc, C, h * w);
        t.p = nhwc;
    } else t.p = const_cast<float *>(src);
    t_in.reset();
    CK(cudaMemsetAsync(ws->d_flag, 0, sizeof(int), st));
    Exec ex{m, ws, st, false};
    ex.amx = amx; ex.final_leaf = final_linear(*pl.root);
    if (amx) amx->done = false;
    ForwardResult r; r.lens = lens0;
    r.y = ex.run(*pl.root, t, r.lens);
    CK(cudaMemcpyAsync(ws->h_flag, ws->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    m->prof_us[1] += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - prof_t1).count();
    return r;
}

// Runs `body` (forward + post-processing of one ABI call).  The tensor-core layers read fp16 operand planes; if any producer saw
// an activation beyond the fp16 range it raised the device flag, and the whole call is repeated on the fp32 CUDA-core kernels
// (still on the GPU: there is no CPU path).  The check needs the stream to be idle, which host-output calls are anyway.
template <class F>
static void run_with_range_fallback(kb_model *m, Workspace *ws, cudaStream_t st, F body) {
    m->force_ffma = false;
    body();
    CK(cudaStreamSynchronize(st));
    if (ws->h_flag && *ws->h_flag) {
        if (getenv("KB_DEBUG")) fprintf(stderr, "[kb] activation outside the fp16 operand range: repeating the call on the fp32 CUDA-core kernels\n");
        ++m->overflow_reruns;
        m->force_ffma = true;
        try { body(); CK(cudaStreamSynchronize(st)); } catch (...) { m->force_ffma = false; throw; }
        m->force_ffma = false;
    }
}

static void collect_timing(Workspace *ws) {
    if (!ws->timing) return;
    for (size_t i = 0; i < ws->n_stages; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ws->stages[i].a, ws->stages[i].b) == cudaSuccess) ws->stages[i].ms = ms;
        else { cudaGetLastError(); ws->stages[i].ms = 0.f; }
    }
}

// NHWC device tensor -> NCHW destination (device or host)
static void emit_nchw(kb_model *m, Workspace *ws, const Tensor &y, float *out, int out_on_device, cudaStream_t st) {
    const size_t elems = (size_t)y.numel();
    if (!elems) return;
    float *dst = out_on_device ? out : (float *)ws->arena.alloc(elems * sizeof(float));
    if (y.c == 1) CK(cudaMemcpyAsync(dst, y.p, elems * sizeof(float), cudaMemcpyDeviceToDevice, st));
    else {
        const int R = (int)(y.h * y.w), Cc = (int)y.c;
        dim3 grid((unsigned)((Cc + 31) / 32), (unsigned)((R + 31) / 32), (unsigned)y.n);
        LAUNCH(m, k_transpose, grid, dim3(32, 8), 0, st, y.p, dst, R, Cc);
    }
    if (!out_on_device) CK(cudaMemcpyAsync(out, dst, elems * sizeof(float), cudaMemcpyDeviceToHost, st));
}

template <typename F>
static int guarded(F &&f) {
    try { return f(); }
    catch (const SpecError &e) { return fail(KB_ERR_SPEC, e.what()); }
    catch (const ShapeError &e) { return fail(KB_ERR_SHAPE, e.what()); }
    catch (const Unsupported &e) { return fail(KB_ERR_UNSUPPORTED, e.what()); }
    catch (const CudaError &e) { cudaGetLastError(); return fail(KB_ERR_CUDA, e.what()); }
    catch (const std::invalid_argument &e) { return fail(KB_ERR_ARG, e.what()); }
    catch (const std::bad_alloc &) { return fail(KB_ERR_STATE, "out of host memory"); }
    catch (const std::exception &e) { return fail(KB_ERR_STATE, e.what()); }
}

static void infer(const Node &n, Dims &d, Lens &l) {
    if (n.kind == K_SERIES) { for (auto &c : n.children) infer(*c, d, l); return; }
    if (n.kind == K_PARALLEL) {
        Dims first; Lens last = l; int64_t ctot = 0; bool have = false;
        for (auto &c : n.children) {
            Dims dc = d; Lens lc = l; infer(*c, dc, lc);
            if (have && (dc.h != first.h || dc.w != first.w)) throw ShapeError("Output shape in parallel block not equal!");
            first = dc; have = true; ctot += dc.c; last = lc;
        }
        d = first; d.c = ctot; l = last; return;
    }
    Dims o = leaf_dims(n, d);
    if (n.kind == K_LSTM && !n.transpose && l.has && d.h != 1)
        throw ShapeError("Height has to be 1 (not " + std::to_string(d.h) + ") for batching/multi-sequences.");
    if (l.has) for (auto &v : l.v) v = leaf_len(n, v, d, o);
    d = o;
}

static size_t decode_bytes(int n, int T, int max_out) {
    return (size_t)n * T * 8 + (size_t)n * max_out * 16 + (size_t)n * 4 + (size_t)n * 24 + 12 * 256;
}

// Fixed-stride result block of a recognition call: [labels | starts | ends | confs] (n x max_out each) + counts (n).
static size_t result_block_bytes(int n, int max_out) { return (size_t)n * max_out * 16 + (size_t)n * 4; }

// Enqueues the CTC collapse and the copy of the result block into the pinned buffer `*pinned` (grown as needed); returns the block
// size.  Nothing here waits for the device.
static size_t decode_enqueue(int n, int T, int max_out, const int *d_lab, const float *d_conf, const int *d_lens, cudaStream_t st,
                             Arena &arena, void **pinned, size_t *pinned_cap, int64_t *launches, RecordXform rx = RecordXform{nullptr, 0, nullptr, nullptr, 0}) {
    const size_t per = (size_t)n * max_out;
    const size_t blk = result_block_bytes(n, max_out);
    char *d = (char *)arena.alloc(blk);
    int *o_lab = (int *)d, *o_start = (int *)(d + per * 4), *o_end = (int *)(d + per * 8);
    float *o_conf = (float *)(d + per * 12); int *o_cnt = (int *)(d + per * 16);
    const int staged = (size_t)T * 8 <= 40 * 1024;           // labels + confidences of one line in shared memory
    k_ctc_collapse<<<(unsigned)n, 256, (size_t)((T + 31) / 32 + 1 + (staged ? 2 * T : 0)) * sizeof(int), st>>>(d_lab, d_conf, d_lens, n, T, max_out, o_lab, o_start, o_end, o_conf, o_cnt, staged, rx);
    ++*launches;
    CK(cudaPeekAtLastError());
    if (blk > *pinned_cap) {
        CK(cudaStreamSynchronize(st));                    // an earlier copy into the old buffer may still be in flight
        if (*pinned) cudaFreeHost(*pinned);
        *pinned = nullptr; *pinned_cap = 0;
        CK(cudaHostAlloc(pinned, blk, cudaHostAllocDefault));
        *pinned_cap = blk;
    }
    CK(cudaMemcpyAsync(*pinned, d, blk, cudaMemcpyDeviceToHost, st));
    return blk;
}

// Result block (host, complete) -> the caller's arrays.  Only the valid prefix of every line is defined in the block; the rest of
// the caller's arrays is zeroed.  counts[i] is the number of labels the line decoded to and may exceed max_out (then only the first
// max_out are stored).
static void decode_unpack(const void *block, int n, int max_out, int32_t *labels, int32_t *starts, int32_t *ends, float *confs, int32_t *counts) {
    const size_t per = (size_t)n * max_out;
    const char *hsrc = (const char *)block;
    const int32_t *h_cnt = (const int32_t *)(hsrc + per * 16);
    for (int i = 0; i < n; ++i) {
        const int c = std::min<int>(h_cnt[i], max_out);
        counts[i] = h_cnt[i];
        const size_t o = (size_t)i * max_out;
        memcpy(labels + o, hsrc + o * 4, (size_t)c * 4); memset(labels + o + c, 0, (size_t)(max_out - c) * 4);
        memcpy(starts + o, hsrc + per * 4 + o * 4, (size_t)c * 4); memset(starts + o + c, 0, (size_t)(max_out - c) * 4);
        memcpy(ends + o, hsrc + per * 8 + o * 4, (size_t)c * 4); memset(ends + o + c, 0, (size_t)(max_out - c) * 4);
        memcpy(confs + o, hsrc + per * 12 + o * 4, (size_t)c * 4); memset(confs + o + c, 0, (size_t)(max_out - c) * 4);
    }
}

static void host_prof_tick(kb_model *m) {
    if (++m->prof_calls % 50 == 0 && getenv("KB_HOST_PROF")) {
        fprintf(stderr, "[kb] host us per call: plan %.1f  launch %.1f  wait %.1f  unpack %.1f\n", m->prof_us[0] / 50, m->prof_us[1] / 50,
                m->prof_us[2] / 50, m->prof_us[3] / 50);
        m->prof_us[0] = m->prof_us[1] = m->prof_us[2] = m->prof_us[3] = 0;
    }
}

// uint8 line images -> the float network input on the device (scale, invert, zero padding; kb_recognize_u8 / async dtype 1)
static const float *stage_u8_lines(kb_model *m, Workspace *ws, const uint8_t *lines, int lines_on_device, int n, int h, int w,
                                   const int32_t *widths, const int16_t *invert_max, cudaStream_t st) {
    const int C = m->plan->input[1];
    const size_t elems = (size_t)n * C * h * w;
    const size_t meta = ((size_t)n * 6 + 15) & ~(size_t)15;                   // widths (int32) + invert_max (int16) behind the pixels
    if (elems + meta + 16 > ws->u8_raw_cap) {
        CK(cudaStreamSynchronize(st));
        if (ws->u8_raw) cudaFree(ws->u8_raw);
        ws->u8_raw = nullptr; ws->u8_raw_cap = 0;
        CK(cudaMalloc((void **)&ws->u8_raw, elems + meta + 16));
        ws->u8_raw_cap = elems + meta + 16;
    }
    if (elems * sizeof(float) > ws->u8_f32_cap) {
        CK(cudaStreamSynchronize(st));
        if (ws->u8_f32) cudaFree(ws->u8_f32);
        ws->u8_f32 = nullptr; ws->u8_f32_cap = 0;
        CK(cudaMalloc((void **)&ws->u8_f32, elems * sizeof(float)));
        ws->u8_f32_cap = elems * sizeof(float);
    }
    const uint8_t *src = lines;
    if (!lines_on_device) { CK(cudaMemcpyAsync(ws->u8_raw, lines, elems, cudaMemcpyHostToDevice, st)); src = ws->u8_raw; }
    uint8_t *mp = ws->u8_raw + ((elems + 15) & ~(size_t)15);
    int *d_w = nullptr; short *d_inv = nullptr;
    if (widths) { d_w = (int *)mp; CK(cudaMemcpyAsync(d_w, widths, (size_t)n * 4, cudaMemcpyHostToDevice, st)); }
    if (invert_max) { d_inv = (short *)(mp + (size_t)n * 4); CK(cudaMemcpyAsync(d_inv, invert_max, (size_t)n * 2, cudaMemcpyHostToDevice, st)); }
    LAUNCH(m, k_u8_lines_to_f32, grid1d((long long)elems, 256, m->sm_count), 256, 0, st, src, ws->u8_f32, (int)n, C, (int)h, (int)w, d_w, d_inv);
    return ws->u8_f32;
}

// One recognition pass on workspace `ws` and stream `st`: net -> softmax statistics -> arg-max -> CTC collapse -> result block on
// its way into ws->pinned.  Enqueue only; the caller synchronises (or records ws->done) and unpacks.
struct RecognizeArgs {
    const float *lines; int lines_on_device; int n, h, w; const int32_t *widths; float temperature; int max_out;
    float *probs; int probs_on_device;
    const int32_t *orig_widths = nullptr; int padding = 0;      // kb_recognize_records: code points + `_scale_val` positions instead of labels / time steps
};
static void recognize_enqueue(kb_model *m, Workspace *ws, const RecognizeArgs &a, cudaStream_t st, int T, int C, std::vector<int32_t> &olens) {
    const int n = a.n;
    size_t extra = decode_bytes(n, T, a.max_out) + (a.probs && !a.probs_on_device ? (size_t)n * C * T * 4 + 4096 : 0);
    // without a probability request the head's arg-max / softmax statistics come out of the final Linear's epilogue (KB_ARGMAX=0:
    // separate kernel over the logits, as with probabilities)
    Exec::ArgmaxOut am; am.temperature = a.temperature;
    const bool fuse_am = !a.probs && !(getenv("KB_ARGMAX") && atoi(getenv("KB_ARGMAX")) == 0);
    ForwardResult r = forward_impl(m, ws, a.lines, a.lines_on_device, n, a.h, a.w, a.widths, st, extra, fuse_am ? &am : nullptr);
    std::unique_ptr<StageTimer> t_dec(new StageTimer(ws, st, "decode", true));
    const long long rows = (long long)n * T;
    int *d_lab = am.done ? am.lab : (int *)ws->arena.alloc((size_t)rows * 4);
    float *d_conf = am.done ? am.conf : (float *)ws->arena.alloc((size_t)rows * 4);
    if (!am.done) LAUNCH(m, k_row_argmax_softmax, (unsigned)((rows + 7) / 8), 256, 0, st, r.y.p, rows, C, a.temperature, d_lab, d_conf);
    olens.resize(n);
    for (int i = 0; i < n; ++i) olens[i] = r.lens.has ? r.lens.v[i] : T;
    int *d_lens = (int *)ws->arena.alloc((size_t)n * 4);
    CK(cudaMemcpyAsync(d_lens, olens.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    if (a.probs) {
        float *dp = a.probs_on_device ? a.probs : (float *)ws->arena.alloc((size_t)n * C * T * 4);
        const size_t smem = (size_t)32 * (C + 1) * 4;
        if (smem <= 48 * 1024) LAUNCH(m, k_probs_nct, dim3((unsigned)((T + 31) / 32), (unsigned)n), 256, smem, st, r.y.p, dp, T, C, a.temperature);
        else LAUNCH(m, k_probs_nct_simple, (unsigned)((rows + 7) / 8), 256, 0, st, r.y.p, dp, n, T, C, a.temperature);
        if (!a.probs_on_device) CK(cudaMemcpyAsync(a.probs, dp, (size_t)n * C * T * 4, cudaMemcpyDeviceToHost, st));
    }
    RecordXform rx{nullptr, 0, nullptr, nullptr, 0};
    if (a.orig_widths) {
        // per line: net_scale = line width / output length, in_scale = original width / (line width - 2 padding)   (rpred.py:143-146)
        std::vector<double> sc((size_t)2 * n); std::vector<int> mv((size_t)n);
        for (int i = 0; i < n; ++i) {
            const int wi = a.widths ? a.widths[i] : a.w;
            if (olens[i] <= 0 || wi - 2 * a.padding == 0) throw ShapeError("record assembly: empty line or line no wider than its padding");
            sc[(size_t)2 * i] = (double)wi / (double)olens[i];
            sc[(size_t)2 * i + 1] = (double)a.orig_widths[i] / (double)(wi - 2 * a.padding);
            mv[(size_t)i] = a.orig_widths[i];
        }
        double *d_sc = (double *)ws->arena.alloc((size_t)2 * n * sizeof(double));
        int *d_mv = (int *)ws->arena.alloc((size_t)n * sizeof(int));
        CK(cudaMemcpyAsync(d_sc, sc.data(), (size_t)2 * n * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_mv, mv.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
        rx = RecordXform{m->codec_lut, m->codec_n, d_sc, d_mv, a.padding};
    }
    decode_enqueue(n, T, a.max_out, d_lab, d_conf, d_lens, st, ws->arena, &ws->pinned, &ws->pinned_cap, &m->launches, rx);
    t_dec.reset();
}

// Forced alignment (kraken/tasks/align.py:111-137): forward -> probabilities (N, C, T) in the arena -> k_forced_align -> segments to the
// caller's host arrays.  `tokens` / `tok_off`: the label sequences, concatenated.  See csrc/align.cuh.
struct AlignArgs {
    const float *lines; int lines_on_device; int n, h, w; const int32_t *widths; float temperature;
    const int32_t *tokens, *tok_off; int jmax, max_seg;
    const int32_t *orig_widths; int padding;
    int32_t *seg_token, *seg_start, *seg_end; float *seg_score; int32_t *seg_counts;
};
static size_t align_bytes(int n, int C, int T, int jmax, int max_seg, int ntok) {
    return (size_t)n * (T + 1) * (jmax + 1) * 4 + (size_t)n * T * 20 + (size_t)n * max_seg * 16 + (size_t)n * 32 + (size_t)(ntok + n + 1) * 4 + 16 * 256;
}
// probabilities (N, C, T) on the device -> segments in the caller's host arrays (enqueue only; the caller synchronises the stream)
static void align_from_probs(const float *dp, const AlignArgs &a, const std::vector<int32_t> &olens, int T, int C, Arena &ar, cudaStream_t st, int64_t *launches) {
    const int n = a.n, ntok = a.tok_off[n];
    const size_t per = (size_t)n * a.max_seg;
    fa::AlignParams p;
    p.probs = dp; p.N = n; p.C = C; p.T = T; p.Jmax = a.jmax; p.max_seg = a.max_seg;
    int *d_lens = (int *)ar.alloc((size_t)n * 4), *d_tok = (int *)ar.alloc((size_t)std::max(ntok, 1) * 4), *d_off = (int *)ar.alloc((size_t)(n + 1) * 4);
    CK(cudaMemcpyAsync(d_lens, olens.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    if (ntok) CK(cudaMemcpyAsync(d_tok, a.tokens, (size_t)ntok * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_off, a.tok_off, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st));
    p.lens = d_lens; p.tokens = d_tok; p.tok_off = d_off;
    p.trellis = (float *)ar.alloc((size_t)n * (T + 1) * (a.jmax + 1) * 4);
    p.fmax = (float *)ar.alloc((size_t)n * T * 4); p.flse = (float *)ar.alloc((size_t)n * T * 4); p.fe0 = (float *)ar.alloc((size_t)n * T * 4);
    p.ptok = (int *)ar.alloc((size_t)n * T * 4); p.pprob = (float *)ar.alloc((size_t)n * T * 4);
    p.seg_token = (int *)ar.alloc(per * 4); p.seg_start = (int *)ar.alloc(per * 4); p.seg_end = (int *)ar.alloc(per * 4);
    p.seg_score = (float *)ar.alloc(per * 4); p.seg_count = (int *)ar.alloc((size_t)n * 4);
    p.scale = nullptr; p.maxv = nullptr; p.padding = a.padding;
    if (a.orig_widths) {                                     // `_scale_val` of the borders, scales as in recognize_enqueue (align.py:128-132)
        std::vector<double> sc((size_t)2 * n); std::vector<int> mv((size_t)n);
        for (int i = 0; i < n; ++i) {
            const int wi = a.widths ? a.widths[i] : a.w;
            if (olens[i] <= 0 || wi - 2 * a.padding == 0) throw ShapeError("forced alignment: empty line or line no wider than its padding");
            sc[(size_t)2 * i] = (double)wi / (double)olens[i];
            sc[(size_t)2 * i + 1] = (double)a.orig_widths[i] / (double)(wi - 2 * a.padding);
            mv[(size_t)i] = a.orig_widths[i];
        }
        double *d_sc = (double *)ar.alloc((size_t)2 * n * sizeof(double));
        int *d_mv = (int *)ar.alloc((size_t)n * sizeof(int));
        CK(cudaMemcpyAsync(d_sc, sc.data(), (size_t)2 * n * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_mv, mv.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
        // (pageable sources: cudaMemcpyAsync returns once they are staged, so the vectors may go out of scope)
        p.scale = d_sc; p.maxv = d_mv;
    }
    CK(cudaMemsetAsync(p.seg_token, 0, per * 4, st)); CK(cudaMemsetAsync(p.seg_start, 0, per * 4, st));
    CK(cudaMemsetAsync(p.seg_end, 0, per * 4, st)); CK(cudaMemsetAsync(p.seg_score, 0, per * 4, st));
    fa::k_forced_align<<<(unsigned)n, 256, fa::align_smem(a.jmax), st>>>(p);
    CK(cudaPeekAtLastError());
    if (launches) ++*launches;
    CK(cudaMemcpyAsync(a.seg_token, p.seg_token, per * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(a.seg_start, p.seg_start, per * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(a.seg_end, p.seg_end, per * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(a.seg_score, p.seg_score, per * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(a.seg_counts, p.seg_count, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
}
static void align_enqueue(kb_model *m, Workspace *ws, const AlignArgs &a, cudaStream_t st, int T, int C, std::vector<int32_t> &olens) {
    const int n = a.n;
    const size_t extra = (size_t)n * C * T * 4 + align_bytes(n, C, T, a.jmax, a.max_seg, a.tok_off[n]);
    ForwardResult r = forward_impl(m, ws, a.lines, a.lines_on_device, n, a.h, a.w, a.widths, st, extra);
    StageTimer t_al(ws, st, "align", true);
    const long long rows = (long long)n * T;
    olens.resize(n);
    for (int i = 0; i < n; ++i) olens[i] = r.lens.has ? r.lens.v[i] : T;
    float *dp = (float *)ws->arena.alloc((size_t)n * C * T * 4);
    const size_t psm = (size_t)32 * (C + 1) * 4;
    if (psm <= 48 * 1024) LAUNCH(m, k_probs_nct, dim3((unsigned)((T + 31) / 32), (unsigned)n), 256, psm, st, r.y.p, dp, T, C, a.temperature);
    else LAUNCH(m, k_probs_nct_simple, (unsigned)((rows + 7) / 8), 256, 0, st, r.y.p, dp, n, T, C, a.temperature);
    align_from_probs(dp, a, olens, T, C, ws->arena, st, &m->launches);
}

// argument checks shared by the two alignment entry points; returns the longest label sequence or throws
static int align_check_tokens(int n, const int32_t *tokens, const int32_t *tok_off, int max_seg) {
    if (tok_off[0] != 0) throw std::invalid_argument("tok_off[0] must be 0");
    int jmax = 0;
    for (int i = 0; i < n; ++i) {
        const int j = tok_off[i + 1] - tok_off[i];
        // an empty transcription: the reference indexes tokens[-1] of an empty tensor (align.py:204) and raises IndexError
        if (j <= 0) throw std::invalid_argument("index -1 is out of bounds for dimension 0 with size 0 (line " + std::to_string(i) + " has no labels)");
        jmax = std::max(jmax, j);
    }
    if (!tokens) throw std::invalid_argument("NULL argument");
    if (max_seg < jmax) throw std::invalid_argument("max_seg must be at least the longest label sequence");
    if (fa::align_smem(jmax) > 48 * 1024) throw Unsupported("label sequences longer than 6000 are not supported");
    return jmax;
}

static void recognition_dims(kb_model *m, int n, int h, int w, int *T, int *C) {
    Dims d; d.n = n; d.c = m->plan->input[1]; d.h = h; d.w = w; Lens l;
    infer(*m->plan->root, d, l);
    if (d.h != 1)
        throw ShapeError("Expected dimension 3 to be 1, actual (" + std::to_string(d.n) + ", " + std::to_string(d.c) + ", " + std::to_string(d.h) + ", " + std::to_string(d.w) + ")");
    *T = (int)d.w; *C = (int)d.c;
}

}  // namespace kb

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

#ifndef KB_SOURCE_HASH
#define KB_SOURCE_HASH "unknown"
#endif
int kb_abi_version(void) { return KB_ABI_VERSION; }
const char *kb_last_error(void) { return g_err.c_str(); }
const char *kb_source_hash(void) { return KB_SOURCE_HASH; }
int kb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int kb_model_create(const char *vgsl_spec, kb_model **out) {
    if (!out) return fail(KB_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (!vgsl_spec) return fail(KB_ERR_SPEC, "vgsl specification argument is missing in args.");
    return guarded([&]() {
        auto m = std::make_unique<kb_model>();
        m->plan = parse_spec(vgsl_spec);
        m->lw.resize(m->plan->leaf_nodes.size());
        for (auto &t : m->plan->tensors) {
            LeafWeights &w = m->lw[t.leaf];
            if ((int)w.host.size() <= t.slot) { w.host.resize(t.slot + 1); w.loaded.resize(t.slot + 1, false); }
        }
        *out = m.release();
        return KB_OK;
    });
}
void kb_model_destroy(kb_model *m) { delete m; }

int kb_model_named_spec(const kb_model *m, char *buf, size_t cap) {
    if (!m) return -fail(KB_ERR_ARG, "model is NULL");
    const std::string &s = m->plan->named_spec;
    if (buf && cap) { size_t k = std::min(cap - 1, s.size()); memcpy(buf, s.data(), k); buf[k] = 0; }
    return (int)s.size();
}
int kb_model_input_shape(const kb_model *m, int32_t shape[4]) {
    if (!m || !shape) return fail(KB_ERR_ARG, "NULL argument");
    for (int i = 0; i < 4; ++i) shape[i] = m->plan->input[i];
    return KB_OK;
}
int kb_model_output_shape(const kb_model *m, int32_t shape[4]) {
    if (!m || !shape) return fail(KB_ERR_ARG, "NULL argument");
    for (int i = 0; i < 4; ++i) shape[i] = m->plan->output[i];
    return KB_OK;
}
int kb_model_num_layers(const kb_model *m) { return m ? (int)m->plan->leaf_nodes.size() : -1; }
int kb_model_layer_info(const kb_model *m, int index, kb_layer_info *info) {
    if (!m || !info) return fail(KB_ERR_ARG, "NULL argument");
    if (index < 0 || index >= (int)m->plan->leaf_nodes.size()) return fail(KB_ERR_ARG, "layer index out of range");
    const Node &n = *m->plan->leaf_nodes[index];
    memset(info, 0, sizeof(*info));
    info->kind = n.kind;
    for (int i = 0; i < 4; ++i) info->out_shape[i] = n.out_shape[i];
    snprintf(info->name, sizeof(info->name), "%s", n.name.c_str());
    snprintf(info->path, sizeof(info->path), "%s", n.path.c_str());
    snprintf(info->block, sizeof(info->block), "%s", n.block.c_str());
    return KB_OK;
}
int kb_model_num_tensors(const kb_model *m) { return m ? (int)m->plan->tensors.size() : -1; }
int kb_model_tensor_info(const kb_model *m, int index, char *name, size_t cap, int64_t shape[4], int32_t *ndim) {
    if (!m) return fail(KB_ERR_ARG, "NULL argument");
    if (index < 0 || index >= (int)m->plan->tensors.size()) return fail(KB_ERR_ARG, "tensor index out of range");
    const TensorDecl &t = m->plan->tensors[index];
    if (name && cap) snprintf(name, cap, "%s", t.name.c_str());
    if (shape) for (size_t i = 0; i < 4; ++i) shape[i] = i < t.shape.size() ? t.shape[i] : 1;
    if (ndim) *ndim = (int)t.shape.size();
    return KB_OK;
}

int kb_model_infer_dims(const kb_model *m, int32_t n, int32_t h, int32_t w, int32_t out_nchw[4]) {
    if (!m || !out_nchw) return fail(KB_ERR_ARG, "NULL argument");
    return guarded([&]() {
        Dims d; d.n = n; d.c = m->plan->input[1]; d.h = h; d.w = w; Lens l;
        infer(*m->plan->root, d, l);
        out_nchw[0] = (int32_t)d.n; out_nchw[1] = (int32_t)d.c; out_nchw[2] = (int32_t)d.h; out_nchw[3] = (int32_t)d.w;
        return KB_OK;
    });
}
int kb_model_infer_lens(const kb_model *m, int32_t n, int32_t h, int32_t w, const int32_t *widths, int32_t *out_lens) {
    if (!m || !widths || !out_lens) return fail(KB_ERR_ARG, "NULL argument");
    return guarded([&]() {
        Dims d; d.n = n; d.c = m->plan->input[1]; d.h = h; d.w = w; Lens l; l.has = true; l.v.assign(widths, widths + n);
        infer(*m->plan->root, d, l);
        for (int i = 0; i < n; ++i) out_lens[i] = l.v[i];
        return KB_OK;
    });
}

int kb_model_load_tensor(kb_model *m, const char *name, const float *data, const int64_t *shape, int32_t ndim) {
    if (!m || !name || !data || !shape) return fail(KB_ERR_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(m->mu);
    for (auto &t : m->plan->tensors) {
        if (t.name != name) continue;
        if ((int)t.shape.size() != ndim) return fail(KB_ERR_ARG, std::string("size mismatch for ") + name + ": wrong rank");
        size_t elems = 1;
        for (int i = 0; i < ndim; ++i) {
            if (t.shape[i] != shape[i]) return fail(KB_ERR_ARG, std::string("size mismatch for ") + name);
            elems *= (size_t)shape[i];
        }
        LeafWeights &w = m->lw[t.leaf];
        w.host[t.slot].assign(data, data + elems);
        w.loaded[t.slot] = true;
        m->finalized = false;
        return KB_OK;
    }
    return fail(KB_ERR_ARG, std::string("unexpected key ") + name);
}

int kb_model_finalize(kb_model *m, int device) {
    if (!m) return fail(KB_ERR_ARG, "model is NULL");
    std::lock_guard<std::mutex> lk(m->mu);
    return guarded([&]() {
        int cnt = 0;
        if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
            cudaGetLastError();
            throw CudaError("no CUDA device available: kraken_b200 has no CPU fallback");
        }
        DeviceGuard dguard0;
        if (device < 0 || device >= cnt) throw CudaError("invalid device ordinal " + std::to_string(device));
        if (m->device >= 0 && m->device != device) {
            CK(cudaSetDevice(m->device));
            m->release_device_state();
        }
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) throw CudaError(std::string("device ") + prop.name + " is not a Blackwell (sm_100) GPU; this library only carries sm_100a code");
        m->sm_count = prop.multiProcessorCount;
        m->device = device;
        set_kernel_attributes();
        finalize_weights(m);
        for (auto &w : m->wss) ensure_workspace(m, w.get());
        m->finalized = true;
        return KB_OK;
    });
}
int kb_model_device(const kb_model *m) { return m && m->finalized ? m->device : -1; }

int kb_forward(kb_model *m, const float *x, int x_on_device, int32_t n, int32_t h, int32_t w, const int32_t *widths,
               float *out, int out_on_device, int32_t *out_lens, void *stream) {
    if (!m || !x || !out) return fail(KB_ERR_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(m->mu);
    return guarded([&]() {
        DeviceGuard dguard;
        ensure_ready(m);
        Workspace *ws = m->ws0();
        ws->timing = m->timing;
        cudaStream_t st = (cudaStream_t)stream;
        Dims d; d.n = n; d.c = m->plan->input[1]; d.h = h; d.w = w; Lens l;
        infer(*m->plan->root, d, l);
        run_with_range_fallback(m, ws, st, [&]() {
            ForwardResult r = forward_impl(m, ws, x, x_on_device, n, h, w, widths, st, out_on_device ? 0 : (size_t)(d.n * d.c * d.h * d.w) * 4 + 4096);
            { StageTimer tt(ws, st, "emit", true); emit_nchw(m, ws, r.y, out, out_on_device, st); }
            if (out_lens) {
                if (r.lens.has) for (int i = 0; i < n; ++i) out_lens[i] = r.lens.v[i];
                else for (int i = 0; i < n; ++i) out_lens[i] = (int32_t)r.y.w;
            }
        });
        collect_timing(ws);
        return KB_OK;
    });
}

}  // extern "C" (reopened below)

// synchronous recognition on workspace 0 and the caller's stream
static int recognize_locked(kb_model *m, const float *lines, int lines_on_device, int32_t n, int32_t h, int32_t w, const int32_t *widths,
                 float temperature, int32_t *labels, int32_t *starts, int32_t *ends, float *confs, int32_t *counts,
                 int32_t max_out, int32_t *out_lens, float *probs, int probs_on_device, void *stream,
                 const int32_t *orig_widths = nullptr, int32_t padding = 0) {
    cudaStream_t st = (cudaStream_t)stream;
    Workspace *ws = m->ws0();
    ws->timing = m->timing;
    int T, C;
    recognition_dims(m, n, h, w, &T, &C);
    std::vector<int32_t> olens;
    RecognizeArgs a{lines, lines_on_device, n, h, w, widths, temperature, max_out, probs, probs_on_device};
    a.orig_widths = orig_widths; a.padding = padding;
    if (orig_widths && !m->codec_lut) {                      // the table is uploaded lazily, on the model's device
        if (m->codec_host.empty()) throw SpecError("kb_recognize_records: call kb_model_set_codec() first");
        CK(cudaMalloc((void **)&m->codec_lut, m->codec_host.size() * sizeof(unsigned)));
        CK(cudaMemcpy(m->codec_lut, m->codec_host.data(), m->codec_host.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
        m->codec_n = (int)m->codec_host.size();
    }
    const auto t0 = std::chrono::steady_clock::now();
    run_with_range_fallback(m, ws, st, [&]() { recognize_enqueue(m, ws, a, st, T, C, olens); });
    const auto t1 = std::chrono::steady_clock::now();
    decode_unpack(ws->pinned, n, max_out, labels, starts, ends, confs, counts);
    m->prof_us[3] += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t1).count();
    m->prof_us[2] += std::chrono::duration<double, std::micro>(t1 - t0).count();
    host_prof_tick(m);
    if (out_lens) for (int i = 0; i < n; ++i) out_lens[i] = olens[i];
    collect_timing(ws);
    return KB_OK;
}

extern "C" {

int kb_recognize(kb_model *m, const float *lines, int lines_on_device, int32_t n, int32_t h, int32_t w, const int32_t *widths,
                 float temperature, int32_t *labels, int32_t *starts, int32_t *ends, float *confs, int32_t *counts,
                 int32_t max_out, int32_t *out_lens, float *probs, int probs_on_device, void *stream) {
    if (!m || !lines || !labels || !starts || !ends || !confs || !counts) return fail(KB_ERR_ARG, "NULL argument");
    if (max_out <= 0) return fail(KB_ERR_ARG, "max_out must be positive");
    if (!(temperature > 0.f)) return fail(KB_ERR_ARG, "temperature must be positive");
    std::lock_guard<std::mutex> lk(m->mu);
    return guarded([&]() {
        DeviceGuard dguard;
        ensure_ready(m);
        return recognize_locked(m, lines, lines_on_device, n, h, w, widths, temperature, labels, starts, ends, confs, counts, max_out,
                                out_lens, probs, probs_on_device, stream);
    });
}

int kb_recognize_u8(kb_model *m, const uint8_t *lines, int lines_on_device, int32_t n, int32_t h, int32_t w, const int32_t *widths,
                    const int16_t *invert_max, float temperature, int32_t *labels, int32_t *starts, int32_t *ends, float *confs,
                    int32_t *counts, int32_t max_out, int32_t *out_lens, float *probs, int probs_on_device, void *stream) {
    if (!m || !lines || !labels || !starts || !ends || !confs || !counts) return fail(KB_ERR_ARG, "NULL argument");
    if (max_out <= 0) return fail(KB_ERR_ARG, "max_out must be positive");
    if (!(temperature > 0.f)) return fail(KB_ERR_ARG, "temperature must be positive");
    if (n <= 0 || h <= 0 || w <= 0) return fail(KB_ERR_SHAPE, "empty input batch");
    std::lock_guard<std::mutex> lk(m->mu);
    return guarded([&]() {
        DeviceGuard dguard;
        ensure_ready(m);
        const float *f32 = stage_u8_lines(m, m->ws0(), lines, lines_on_device, n, h, w, widths, invert_max, (cudaStream_t)stream);
        return recognize_locked(m, f32, 1, n, h, w, widths, temperature, labels, starts, ends, confs, counts, max_out, out_lens,
                                probs, probs_on_device, stream);
    });
}

int kb_model_set_codec(kb_model *m, const uint32_t *l2c, int32_t n_labels) {
    if (!m || (n_labels > 0 && !l2c) || n_labels < 0) return fail(KB_ERR_ARG, "invalid codec table");
    std::lock_guard<std::mutex> lk(m->mu);
    return guarded([&]() {
        m->codec_host.assign(l2c, l2c + n_labels);
        if (m->codec_lut) {
            DeviceGuard dguard;
            CK(cudaSetDevice(m->device));
            CK(cudaDeviceSynchronize());
            cudaFree(m->codec_lut);
            m->codec_lut = nullptr; m->codec_n = 0;
        }
        return (int)KB_OK;
    });
}

int kb_recognize_records(kb_model *m, const void *lines, int dtype, int lines_on_device, int32_t n, int32_t h, int32_t w,
                         const int32_t *widths, const int16_t *invert_max, float temperature, const int32_t *orig_widths, int32_t padding,
                         uint32_t *codepoints, int32_t *starts, int32_t *ends, float *confs, int32_t *counts, int32_t max_out,
                         int32_t *out_lens, void *stream) {
    if (!m || !lines || !orig_widths || !codepoints || !starts || !ends || !confs || !counts) return fail(KB_ERR_ARG, "NULL argument");
    if (dtype != KB_DTYPE_F32 && dtype != KB_DTYPE_U8) return fail(KB_ERR_ARG, "dtype must be KB_DTYPE_F32 or KB_DTYPE_U8");
    if (max_out <= 0 || padding < 0) return fail(KB_ERR_ARG, "max_out must be positive, padding non-negative");
    if (!(temperature > 0.f)) return fail(KB_ERR_ARG, "temperature must be positive");
    if (n <= 0 || h <= 0 || w <= 0) return fail(KB_ERR_SHAPE, "empty input batch");
    std::lock_guard<std::mutex> lk(m->mu);
    return guarded([&]() {
        DeviceGuard dguard;
        ensure_ready(m);
        const float *f32 = (const float *)lines; int on_dev = lines_on_device;
        if (dtype == KB_DTYPE_U8) { f32 = stage_u8_lines(m, m->ws0(), (const uint8_t *)lines, lines_on_device, n, h, w, widths, invert_max, (cudaStream_t)stream); on_dev = 1; }
        return recognize_locked(m, f32, on_dev, n, h, w, widths, temperature, (int32_t *)codepoints, starts, ends, confs, counts, max_out, out_lens,
                                nullptr, 0, stream, orig_widths, padding);
    });
}

/* ---- forced alignment ----------------------------------------------------------------------------------------------------------- */
int kb_forced_align(kb_model *m, const void *lines, int dtype, int lines_on_device, int32_t n, int32_t h, int32_t w,
                    const int32_t *widths, const int16_t *invert_max, float temperature, const int32_t *tokens, const int32_t *tok_off,
                    const int32_t *orig_widths, int32_t padding, int32_t *seg_token, int32_t *seg_start, int32_t *seg_end,
                    float *seg_score, int32_t *seg_counts, int32_t max_seg, int32_t *out_lens, void *stream) {
    if (!m || !lines || !tok_off || !seg_token || !seg_start || !seg_end || !seg_score || !seg_counts) return fail(KB_ERR_ARG, "NULL argument");
    if (dtype != KB_DTYPE_F32 && dtype != KB_DTYPE_U8) return fail(KB_ERR_ARG, "dtype must be KB_DTYPE_F32 or KB_DTYPE_U8");
    if (!(temperature > 0.f)) return fail(KB_ERR_ARG, "temperature must be positive");
    if (n <= 0 || h <= 0 || w <= 0) return fail(KB_ERR_SHAPE, "empty input batch");
    if (padding < 0) return fail(KB_ERR_ARG, "padding must be non-negative");
    std::lock_guard<std::mutex> lk(m->mu);
    return guarded([&]() {
        DeviceGuard dguard;
        ensure_ready(m);
        cudaStream_t st = (cudaStream_t)stream;
        Workspace *ws = m->ws0();
        ws->timing = m->timing;
        const float *f32 = (const float *)lines; int on_dev = lines_on_device;
        if (dtype == KB_DTYPE_U8) { f32 = stage_u8_lines(m, ws, (const uint8_t *)lines, lines_on_device, n, h, w, widths, invert_max, st); on_dev = 1; }
        const int jmax = align_check_tokens(n, tokens, tok_off, max_seg);
        int T, C;
        recognition_dims(m, n, h, w, &T, &C);
        for (int i = 0; i < tok_off[n]; ++i)
            if (tokens[i] < 0 || tokens[i] >= C) throw ShapeError("label " + std::to_string(tokens[i]) + " outside the model's " + std::to_string(C) + " classes");
        std::vector<int32_t> olens;
        AlignArgs a{f32, on_dev, n, h, w, widths, temperature, tokens, tok_off, jmax, max_seg, orig_widths, padding,
                    seg_token, seg_start, seg_end, seg_score, seg_counts};
        run_with_range_fallback(m, ws, st, [&]() { align_enqueue(m, ws, a, st, T, C, olens); });
        if (out_lens) for (int i = 0; i < n; ++i) out_lens[i] = olens[i];
        collect_timing(ws);
        return (int)KB_OK;
    });
}

int kb_forced_align_probs(const float *probs, int probs_on_device, int32_t n, int32_t c, int32_t t, const int32_t *lens,
                          const int32_t *tokens, const int32_t *tok_off, int32_t *seg_token, int32_t *seg_start, int32_t *seg_end,
                          float *seg_score, int32_t *seg_counts, int32_t max_seg, int device, void *stream) {
    if (!probs || !tok_off || !seg_token || !seg_start || !seg_end || !seg_score || !seg_counts) return fail(KB_ERR_ARG, "NULL argument");
    if (n <= 0 || c <= 0 || t <= 0) return fail(KB_ERR_ARG, "invalid sizes");
    return guarded([&]() {
        const int jmax = align_check_tokens(n, tokens, tok_off, max_seg);
        for (int i = 0; i < tok_off[n]; ++i)
            if (tokens[i] < 0 || tokens[i] >= c) throw ShapeError("label " + std::to_string(tokens[i]) + " outside the " + std::to_string(c) + " classes");
        int cnt = 0;
        if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) { cudaGetLastError(); throw CudaError("no CUDA device available: kraken_b200 has no CPU fallback"); }
        DeviceGuard dguard;
        CK(cudaSetDevice(device));
        cudaStream_t st = (cudaStream_t)stream;
        std::vector<int32_t> olens((size_t)n);
        for (int i = 0; i < n; ++i) olens[i] = lens ? std::min(std::max(lens[i], 0), t) : t;
        Arena ar;
        const size_t in_bytes = (size_t)n * c * t * 4;
        const size_t total = (probs_on_device ? 0 : in_bytes + 256) + align_bytes(n, c, t, jmax, max_seg, tok_off[n]) + 8192;
        CK(cudaMalloc((void **)&ar.base, total)); ar.cap = total;
        try {
            const float *dp = probs;
            if (!probs_on_device) { float *b = (float *)ar.alloc(in_bytes); CK(cudaMemcpyAsync(b, probs, in_bytes, cudaMemcpyHostToDevice, st)); dp = b; }
            AlignArgs a{nullptr, 0, n, 1, t, nullptr, 1.f, tokens, tok_off, jmax, max_seg, nullptr, 0, seg_token, seg_start, seg_end, seg_score, seg_counts};
            align_from_probs(dp, a, olens, t, c, ar, st, nullptr);
            CK(cudaStreamSynchronize(st));
        } catch (...) { cudaFree(ar.base); throw; }
        cudaFree(ar.base);
        return (int)KB_OK;
    });
}

/* ---- asynchronous pipeline ------------------------------------------------------------------------------------------------------ */
int kb_set_pipeline_depth(kb_model *m, int32_t depth) {
    if (!m) return fail(KB_ERR_ARG, "model is NULL");
    if (depth < 1 || depth > 16) return fail(KB_ERR_ARG, "pipeline depth must be in 1..16");
    std::lock_guard<std::mutex> lk(m->mu);
    return guarded([&]() {
        DeviceGuard dguard;
        ensure_ready(m);
        for (size_t i = 1; i < m->wss.size(); ++i)
            if (m->wss[i]->busy) throw SpecError("kb_set_pipeline_depth: tickets are still in flight");
        while ((int)m->wss.size() > depth + 1) { m->wss.back()->release(); m->wss.pop_back(); }
        while ((int)m->wss.size() < depth + 1) {
            m->wss.emplace_back(new Workspace());
            m->wss.back()->index = (int)m->wss.size() - 1;
            ensure_workspace(m, m->wss.back().get());
        }
        return (int)KB_OK;
    });
}
int kb_pipeline_depth(const kb_model *m) { return m ? (int)m->wss.size() - 1 : -1; }

int kb_recognize_async(kb_model *m, const void *lines, int dtype, int lines_on_device, int32_t n, int32_t h, int32_t w, const int32_t *widths,
                       const int16_t *invert_max, float temperature, int32_t max_out, void *input_stream, int64_t *ticket) {
    if (!m || !lines || !ticket) return fail(KB_ERR_ARG, "NULL argument");
    if (dtype != KB_DTYPE_F32 && dtype != KB_DTYPE_U8) return fail(KB_ERR_ARG, "dtype must be KB_DTYPE_F32 or KB_DTYPE_U8");
    if (max_out <= 0) return fail(KB_ERR_ARG, "max_out must be positive");
    if (!(temperature > 0.f)) return fail(KB_ERR_ARG, "temperature must be positive");
    if (n <= 0 || h <= 0 || w <= 0) return fail(KB_ERR_SHAPE, "empty input batch");
    std::lock_guard<std::mutex> lk(m->mu);
    return guarded([&]() {
        DeviceGuard dguard;
        ensure_ready(m);
        if (m->wss.size() < 2) throw SpecError("kb_recognize_async: call kb_set_pipeline_depth() first");
        const int depth = (int)m->wss.size() - 1;
        Workspace *ws = m->wss[1 + (size_t)(m->next_ticket % depth)].get();
        if (ws->busy) throw SpecError("kb_recognize_async: all " + std::to_string(depth) + " pipeline slots hold un-waited tickets (kb_wait the oldest first)");
        ws->timing = false;
        cudaStream_t st = ws->stream;
        if (lines_on_device) {                            // the input was produced on the caller's stream
            CK(cudaEventRecord(ws->input_ready, (cudaStream_t)input_stream));
            CK(cudaStreamWaitEvent(st, ws->input_ready, 0));
        }
        int T, C;
        recognition_dims(m, n, h, w, &T, &C);
        Workspace::Pending &pd = ws->pend;
        pd.n = n; pd.h = h; pd.w = w; pd.T = T; pd.max_out = max_out; pd.dtype = dtype; pd.temperature = temperature;
        pd.has_widths = widths != nullptr;
        if (widths) pd.widths.assign(widths, widths + n); else pd.widths.clear();
        m->force_ffma = false;
        const float *f32 = (const float *)lines; int on_dev = lines_on_device;
        if (dtype == KB_DTYPE_U8) { f32 = stage_u8_lines(m, ws, (const uint8_t *)lines, lines_on_device, n, h, w, widths, invert_max, st); on_dev = 1; }
        pd.device_input = on_dev ? f32 : nullptr;
        RecognizeArgs a{f32, on_dev, n, h, w, widths, temperature, max_out, nullptr, 0};
        recognize_enqueue(m, ws, a, st, T, C, pd.olens);
        if (!on_dev) pd.device_input = ws->arena.base;           // host lines were staged at the start of the slot's arena
        CK(cudaEventRecord(ws->done, st));
        ws->busy = true; ws->ticket = m->next_ticket;
        *ticket = m->next_ticket++;
        return (int)KB_OK;
    });
}

int kb_wait(kb_model *m, int64_t ticket, int32_t *labels, int32_t *starts, int32_t *ends, float *confs, int32_t *counts, int32_t *out_lens) {
    if (!m || !labels || !starts || !ends || !confs || !counts) return fail(KB_ERR_ARG, "NULL argument");
    Workspace *ws = nullptr;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        for (size_t i = 1; i < m->wss.size(); ++i)
            if (m->wss[i]->busy && m->wss[i]->ticket == ticket) ws = m->wss[i].get();
    }
    if (!ws) return fail(KB_ERR_ARG, "kb_wait: unknown or already collected ticket");
    return guarded([&]() {
        DeviceGuard dguard;
        CK(cudaSetDevice(m->device));
        CK(cudaEventSynchronize(ws->done));               // blocking-sync event: the host thread sleeps, the model stays unlocked
        std::lock_guard<std::mutex> lk(m->mu);
        Workspace::Pending &pd = ws->pend;
        try {
            if (*ws->h_flag) {
                // an activation left the fp16 operand range: repeat this batch on the fp32 CUDA-core kernels (input still staged on the device)
                ++m->overflow_reruns;
                m->force_ffma = true;
                int T, C;
                recognition_dims(m, pd.n, pd.h, pd.w, &T, &C);
                RecognizeArgs a{(const float *)pd.device_input, 1, pd.n, pd.h, pd.w, pd.has_widths ? pd.widths.data() : nullptr, pd.temperature, pd.max_out, nullptr, 0};
                // the staged input sits inside the arena this pass is about to reuse: move it out of the way first
                const size_t in_bytes = (size_t)pd.n * m->plan->input[1] * pd.h * pd.w * sizeof(float);
                float *keep = nullptr;
                CK(cudaMalloc((void **)&keep, in_bytes));
                try {
                    CK(cudaMemcpyAsync(keep, pd.device_input, in_bytes, cudaMemcpyDeviceToDevice, ws->stream));
                    a.lines = keep;
                    recognize_enqueue(m, ws, a, ws->stream, T, C, pd.olens);
                    CK(cudaStreamSynchronize(ws->stream));
                } catch (...) { cudaFree(keep); throw; }
                cudaFree(keep);
                m->force_ffma = false;
            }
        } catch (...) { m->force_ffma = false; ws->busy = false; throw; }
        decode_unpack(ws->pinned, pd.n, pd.max_out, labels, starts, ends, confs, counts);
        if (out_lens) for (int i = 0; i < pd.n; ++i) out_lens[i] = pd.olens[i];
        ws->busy = false;
        return (int)KB_OK;
    });
}

int kb_ctc_greedy_decode(const float *probs, int probs_on_device, int32_t n, int32_t c, int32_t w, const int32_t *lens,
                         int32_t *labels, int32_t *starts, int32_t *ends, float *confs, int32_t *counts, int32_t max_out,
                         int device, void *stream) {
    if (!probs || !labels || !starts || !ends || !confs || !counts) return fail(KB_ERR_ARG, "NULL argument");
    if (n <= 0 || c <= 0 || w < 0 || max_out <= 0) return fail(KB_ERR_ARG, "invalid sizes");
    if (!lens && n != 1) return fail(KB_ERR_ARG, "seq_lens need to be set for batch decoding.");    // ctc_decoder.py:60-61
    return guarded([&]() {
        int cnt = 0;
        if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) { cudaGetLastError(); throw CudaError("no CUDA device available: kraken_b200 has no CPU fallback"); }
        DeviceGuard dguard;
        CK(cudaSetDevice(device));
        cudaStream_t st = (cudaStream_t)stream;
        if (w == 0) { for (int i = 0; i < n; ++i) counts[i] = 0; return (int)KB_OK; }
        Arena ar;
        const size_t in_bytes = (size_t)n * c * w * 4;
        size_t total = (probs_on_device ? 0 : in_bytes) + decode_bytes(n, w, max_out) + 8192;
        CK(cudaMalloc((void **)&ar.base, total)); ar.cap = total;
        void *pinned = nullptr; size_t pcap = 0; int64_t launches = 0;
        try {
            const float *dp = probs;
            if (!probs_on_device) { float *t = (float *)ar.alloc(in_bytes); CK(cudaMemcpyAsync(t, probs, in_bytes, cudaMemcpyHostToDevice, st)); dp = t; }
            int *d_lab = (int *)ar.alloc((size_t)n * w * 4); float *d_conf = (float *)ar.alloc((size_t)n * w * 4);
            int *d_lens = nullptr;
            if (lens) { d_lens = (int *)ar.alloc((size_t)n * 4); CK(cudaMemcpyAsync(d_lens, lens, (size_t)n * 4, cudaMemcpyHostToDevice, st)); }
            k_col_argmax<<<(unsigned)(((long long)n * w + 255) / 256), 256, 0, st>>>(dp, n, c, w, d_lab, d_conf);
            CK(cudaPeekAtLastError());
            decode_enqueue(n, w, max_out, d_lab, d_conf, d_lens, st, ar, &pinned, &pcap, &launches);
            CK(cudaStreamSynchronize(st));
            decode_unpack(pinned, n, max_out, labels, starts, ends, confs, counts);
        } catch (...) { cudaFree(ar.base); if (pinned) cudaFreeHost(pinned); throw; }
        cudaFree(ar.base); if (pinned) cudaFreeHost(pinned);
        return (int)KB_OK;
    });
}

int32_t kb_line_width(int32_t box_w, int32_t box_h, int32_t out_h, int32_t pad) {
    const int ow = lp::resized_width(box_w, box_h, out_h);
    return ow < 1 ? 0 : ow + 2 * std::max(pad, 0);
}

int kb_debug_axis_coeffs(int32_t in_size, int32_t out_size, int32_t *ksize, int32_t *bounds, int32_t *kk, int32_t kk_cap) {
    if (in_size <= 0 || out_size <= 0 || !ksize) return fail(KB_ERR_ARG, "invalid sizes");
    const int ks = in_size == out_size ? lp::identity_ksize() : lp::axis_ksize(in_size, out_size);
    *ksize = ks;
    if (!bounds || !kk) return KB_OK;
    if ((long long)kk_cap < (long long)out_size * ks) return fail(KB_ERR_ARG, "kk_cap too small");
    lp::axis_coeffs(in_size, out_size, ks, bounds, kk);
    return KB_OK;
}

int kb_prepare_lines_u8(kb_model *m, const uint8_t *page, int page_on_device, int32_t page_h, int32_t page_w, int32_t channels,
                        int32_t n, const int32_t *boxes, int32_t out_h, int32_t pad, uint8_t *lines, int32_t wmax,
                        int32_t *widths, int16_t *invert_max, void *stream) {
    if (!m || !page || !boxes || !lines || !widths) return fail(KB_ERR_ARG, "NULL argument");
    if (channels != 1 && channels != 3) return fail(KB_ERR_ARG, "page must have 1 ('L') or 3 ('RGB', interleaved) channels");
    if (n <= 0 || page_h <= 0 || page_w <= 0 || out_h <= 0 || pad < 0 || wmax <= 0) return fail(KB_ERR_ARG, "invalid sizes");
    for (int i = 0; i < n; ++i) {
        const int32_t *b = boxes + 4 * i;
        if (b[0] < 0 || b[1] < 0 || b[2] > page_w || b[3] > page_h || b[2] <= b[0] || b[3] <= b[1])
            return fail(KB_ERR_ARG, "Line outside of image bounds");                                   // segmentation.py:1639-1642
        const int wd = kb_line_width(b[2] - b[0], b[3] - b[1], out_h, pad);
        if (wd - 2 * pad < 1) return fail(KB_ERR_SHAPE, "height and width must be > 0");               // what Image.resize raises for such a line
        if (wd > wmax) return fail(KB_ERR_ARG, "wmax is smaller than a prepared line (size it with kb_line_width)");
        widths[i] = wd;
    }
    std::lock_guard<std::mutex> lk(m->mu);
    return guarded([&]() {
        DeviceGuard dguard;
        ensure_ready(m);
        cudaStream_t st = (cudaStream_t)stream;
        lp::Plan pl;
        lp::build_plan(pl, n, boxes, out_h, pad);
        const size_t page_bytes = page_on_device ? 0 : (size_t)page_h * page_w * channels;
        const size_t meta_bytes = (size_t)n * sizeof(lp::LineMeta), tab_bytes = pl.tab.size() * sizeof(int32_t);
        auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
        const size_t host_need = up(meta_bytes) + up(tab_bytes);
        const size_t dev_need = up(page_bytes) + host_need + up(pl.tmp_bytes) + up((size_t)n * sizeof(int));
        if (host_need > m->prep_host_cap || dev_need > m->prep_dev_cap) CK(cudaStreamSynchronize(st));   // an earlier call may still read the old buffers
        if (host_need > m->prep_host_cap) {
            if (m->prep_host) cudaFreeHost(m->prep_host);
            m->prep_host = nullptr; m->prep_host_cap = 0;
            CK(cudaHostAlloc((void **)&m->prep_host, host_need + host_need / 2, cudaHostAllocDefault));
            m->prep_host_cap = host_need + host_need / 2;
        }
        if (dev_need > m->prep_dev_cap) {
            if (m->prep_dev) cudaFree(m->prep_dev);
            m->prep_dev = nullptr; m->prep_dev_cap = 0;
            CK(cudaMalloc((void **)&m->prep_dev, dev_need + dev_need / 2));
            m->prep_dev_cap = dev_need + dev_need / 2;
        }
        char *d_page = m->prep_dev, *d_tabs = d_page + up(page_bytes), *d_tmp = d_tabs + host_need, *d_max = d_tmp + up(pl.tmp_bytes);
        memcpy(m->prep_host, pl.meta.data(), meta_bytes);
        memcpy(m->prep_host + up(meta_bytes), pl.tab.data(), tab_bytes);
        if (!page_on_device) CK(cudaMemcpyAsync(d_page, page, page_bytes, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_tabs, m->prep_host, host_need, cudaMemcpyHostToDevice, st));
        const bool want_max = invert_max && pad == 0;         // with white padding the maximum is 255 by construction
        if (want_max) CK(cudaMemsetAsync(d_max, 0, (size_t)n * sizeof(int), st));
        const uint8_t *pg = page_on_device ? page : (const uint8_t *)d_page;
        const lp::LineMeta *d_meta = (const lp::LineMeta *)d_tabs; const int32_t *d_tab = (const int32_t *)(d_tabs + up(meta_bytes));
        LAUNCH(m, lp::k_prep_horizontal, dim3((unsigned)((pl.max_ow + 127) / 128), (unsigned)pl.max_rows, (unsigned)n), 128, 0, st, pg, (int)page_w,
               (int)channels, d_meta, d_tab, (uint8_t *)d_tmp);
        LAUNCH(m, lp::k_prep_vertical, dim3((unsigned)((pl.max_width + 127) / 128), (unsigned)out_h, (unsigned)n), 128, 0, st, (const uint8_t *)d_tmp,
               d_meta, d_tab, lines, (int)out_h, (int)wmax, (int)pad, want_max ? (int *)d_max : (int *)nullptr);
        std::vector<int> mx((size_t)(want_max ? n : 0));
        if (want_max) CK(cudaMemcpyAsync(mx.data(), d_max, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));                        // lines complete; the pinned tables may be overwritten by the next call
        if (invert_max) for (int i = 0; i < n; ++i) invert_max[i] = want_max ? (int16_t)mx[(size_t)i] : (int16_t)255;
        return (int)KB_OK;
    });
}

int kb_segment(kb_model *m, const float *pages, int pages_on_device, int32_t n, int32_t h, int32_t w, int32_t out_h, int32_t out_w,
               float *heatmap, int heatmap_on_device, void *stream) {
    if (!m || !pages || !heatmap) return fail(KB_ERR_ARG, "NULL argument");
    if (out_h <= 0 || out_w <= 0) return fail(KB_ERR_ARG, "invalid output size");
    std::lock_guard<std::mutex> lk(m->mu);
    return guarded([&]() {
        DeviceGuard dguard;
        ensure_ready(m);
        cudaStream_t st = (cudaStream_t)stream;
        Dims d; d.n = n; d.c = m->plan->input[1]; d.h = h; d.w = w; Lens l;
        infer(*m->plan->root, d, l);
        const size_t out_elems = (size_t)n * d.c * out_h * out_w;
        Workspace *ws = m->ws0();
        ws->timing = m->timing;
        run_with_range_fallback(m, ws, st, [&]() {
            ForwardResult r = forward_impl(m, ws, pages, pages_on_device, n, h, w, nullptr, st, heatmap_on_device ? 0 : out_elems * 4 + 4096);
            std::unique_ptr<StageTimer> t_up(new StageTimer(ws, st, "upsample_sigmoid", true));
            float *dst = heatmap_on_device ? heatmap : (float *)ws->arena.alloc(out_elems * 4);
            LAUNCH(m, k_upsample_sigmoid, grid1d((long long)n * out_h * out_w, 256, m->sm_count), 256, 0, st, r.y.p, dst, (int)r.y.n, (int)r.y.h,
                   (int)r.y.w, (int)r.y.c, out_h, out_w);
            if (!heatmap_on_device) CK(cudaMemcpyAsync(heatmap, dst, out_elems * 4, cudaMemcpyDeviceToHost, st));
            t_up.reset();
        });
        collect_timing(ws);
        return KB_OK;
    });
}

int kb_debug_layer_output(kb_model *m, const char *name, int32_t dims[4], float *out_host, int dims_only) {
    if (!m || !name || !dims) return fail(KB_ERR_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(m->mu);
    return guarded([&]() {
        DeviceGuard dguard;
        ensure_ready(m);
        Workspace *ws = m->ws0();
        auto it = ws->taps.find(name);
        if (it == ws->taps.end()) throw SpecError(std::string("no output recorded for layer ") + name);
        const Tensor &t = it->second;
        if (!t.p && !dims_only) throw SpecError(std::string("layer ") + name + " did not materialise in the last call (fused into its consumer); set KB_FUSE=0 / KB_KEEP_FP32=1, or use kb_forward");
        dims[0] = (int32_t)t.n; dims[1] = (int32_t)t.c; dims[2] = (int32_t)t.h; dims[3] = (int32_t)t.w;
        if (dims_only || !out_host) return (int)KB_OK;
        const size_t elems = (size_t)t.numel();
        if (!elems) return (int)KB_OK;
        float *tmp = nullptr;
        CK(cudaMalloc(&tmp, elems * 4));
        cudaError_t e = cudaSuccess;
        if (t.c == 1) e = cudaMemcpy(tmp, t.p, elems * 4, cudaMemcpyDeviceToDevice);
        else {
            const int R = (int)(t.h * t.w), Cc = (int)t.c;
            dim3 grid((unsigned)((Cc + 31) / 32), (unsigned)((R + 31) / 32), (unsigned)t.n);
            k_transpose<<<grid, dim3(32, 8)>>>(t.p, tmp, R, Cc);
            e = cudaDeviceSynchronize();
        }
        if (e == cudaSuccess) e = cudaMemcpy(out_host, tmp, elems * 4, cudaMemcpyDeviceToHost);
        cudaFree(tmp);
        if (e != cudaSuccess) throw CudaError(std::string("layer output copy failed: ") + cudaGetErrorString(e));
        return (int)KB_OK;
    });
}

int kb_debug_gemm(const float *a, const float *b, const float *bias, float *c, int32_t M, int32_t N, int32_t K, int use_tc, int device) {
    if (!a || !b || !c || M <= 0 || N <= 0 || K <= 0) return fail(KB_ERR_ARG, "invalid arguments");
    return guarded([&]() {
        int cnt = 0;
        if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) { cudaGetLastError(); throw CudaError("no CUDA device available: kraken_b200 has no CPU fallback"); }
        DeviceGuard dguard;
        CK(cudaSetDevice(device));
        kb_model tmp; tmp.device = device; tmp.use_tc = use_tc != 0;
        cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device)); tmp.sm_count = prop.multiProcessorCount;
        LeafWeights w; w.K = K; w.ncols = N; w.ncp = (N + 63) / 64 * 64;
        std::vector<float> rows(b, b + (size_t)N * K), wt((size_t)K * w.ncp, 0.f);
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) wt[(size_t)k * w.ncp + n] = rows[(size_t)n * K + k];
        w.wt = upload(&tmp, wt);
        if (bias) w.bias = upload(&tmp, std::vector<float>(bias, bias + N));
        upload_split(&tmp, rows, w);
        set_kernel_attributes();
        Workspace *ws = tmp.ws0();
        ensure_workspace(&tmp, ws);
        const size_t abytes = (size_t)M * K * 4, cbytes = (size_t)M * N * 4;
        CK(cudaMalloc((void **)&ws->arena.base, 3 * abytes + cbytes + 65536)); ws->arena.cap = 3 * abytes + cbytes + 65536;
        Tensor x; x.n = 1; x.h = 1; x.w = M; x.c = K; x.p = (float *)ws->arena.alloc(abytes);
        float *dc = (float *)ws->arena.alloc(cbytes);
        CK(cudaMemcpy(x.p, a, abytes, cudaMemcpyHostToDevice));
        Exec ex{&tmp, ws, nullptr, false};
        if (use_tc && !ex.tc_eligible(x, w, nullptr)) throw Unsupported("shape not eligible for the tcgen05 GEMM (need K % 8 == 0, K >= 32, N >= 64, M >= 128)");
        ex.gemm(x, w, nullptr, ACT_LINEAR, dc, 1, M);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(c, dc, cbytes, cudaMemcpyDeviceToHost));
        return (int)KB_OK;
    });
}

int64_t kb_launch_count(const kb_model *m) { return m ? m->launches : 0; }
void kb_reset_launch_count(kb_model *m) { if (m) m->launches = 0; }
int64_t kb_range_fallback_count(const kb_model *m) { return m ? m->overflow_reruns : 0; }
int kb_set_timing(kb_model *m, int enabled) { if (!m) return fail(KB_ERR_ARG, "model is NULL"); m->timing = enabled != 0; return KB_OK; }
int kb_timing_count(kb_model *m) {
    if (!m) return -1;
    std::lock_guard<std::mutex> lk(m->mu);
    Workspace *ws = m->ws0();
    return m->timing ? (int)ws->n_stages : 0;
}
int kb_timing_entry(kb_model *m, int index, char *name, size_t cap, float *ms) {
    if (!m) return fail(KB_ERR_ARG, "model is NULL");
    std::lock_guard<std::mutex> lk(m->mu);
    Workspace *ws = m->ws0();
    if (index < 0 || index >= (int)ws->n_stages) return fail(KB_ERR_ARG, "timing index out of range");
    if (name && cap) snprintf(name, cap, "%s", ws->stages[index].name.c_str());
    if (ms) *ms = ws->stages[index].ms;
    return KB_OK;
}

}  // extern "C"
