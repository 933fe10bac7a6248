// vgsl_plan.hpp - VGSL spec front end of the engine: tokenizer/grammar, auto-naming, static and
// runtime shape arithmetic, seq_len propagation.  Pure host C++17 (no CUDA), header-only.
//
// Behavioural contract = TorchVGSLModel.__init__/_parse/build_* of the reference
// (kraken/lib/vgsl/model.py:109-243, 570-902) and the get_shape()/seq_len formulas of
// kraken/lib/vgsl/layers.py.  It is a fresh implementation: a cursor-based prefix matcher per block
// type instead of the reference's regexes, producing a tree of plain structs that the CUDA executor
// walks.  Quirks that are observable through `named_spec`/`output` and therefore reproduced:
//   * every block is matched as a PREFIX; trailing characters are ignored but kept in the named spec
//   * block types are tried in a fixed order (model.py:167-171)
//   * auto names are <type>_<idx>, inserted after the leading non-digit run (model.py:53-64);
//     heat-map outputs are named after their non-linearity letter (model.py:811)
//   * Reshape's static shape is computed on a probe with variable dims set to 1 (layers.py:337-341)
//   * direction 'r' builds a forward LSTM (layers.py:496,507-511), 'G' builds an LSTM too
#pragma once
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace kb {

struct SpecError : std::runtime_error { using std::runtime_error::runtime_error; };       // -> ValueError
struct ShapeError : std::runtime_error { using std::runtime_error::runtime_error; };      // -> KrakenInputException
struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };

enum Kind { K_CONV = 1, K_POOL = 2, K_RESHAPE = 3, K_LSTM = 4, K_DROPOUT = 5, K_GN = 6, K_LINEAR = 7,
            K_ADD = 8, K_IDENTITY = 9, K_SERIES = 100, K_PARALLEL = 101 };
enum Act { ACT_LINEAR = 0, ACT_SIGMOID_LOGITS = 1, ACT_TANH = 2, ACT_SOFTMAX = 3, ACT_RELU = 4, ACT_LEAKY = 5 };

struct Node {
    int kind = 0;
    std::string name, path, block;
    int in_shape[4] = {0, 0, 0, 0}, out_shape[4] = {0, 0, 0, 0};   // static (batch, C, H, W); 0 = variable
    int leaf_index = -1;
    // conv / pool
    int kh = 1, kw = 1, sy = 1, sx = 1, dy = 1, dx = 1, py = 0, px = 0, cin = 0, cout = 0, act = ACT_LINEAR;
    // lstm
    int hidden = 0, bidi = 0, transpose = 0, summarize = 0, legacy = 0;
    // dropout
    double drop_p = 0.5; int drop_dim = 1;
    // groupnorm
    int groups = 0;
    // linear
    int aug = 0;
    // addition
    int add_dim = 0, add_chunk = 0;
    // reshape (NCHW dims)
    int rs_src = 0, rs_a = 0, rs_b = 0, rs_high = 0, rs_low = 0;
    std::vector<std::unique_ptr<Node>> children;
};

struct TensorDecl { std::string name; std::vector<int64_t> shape; int leaf = -1; int slot = 0; };

// ---------------------------------------------------------------------------------------------
// small cursor helpers
// ---------------------------------------------------------------------------------------------
struct Cur {
    const std::string &s; size_t i = 0;
    explicit Cur(const std::string &str) : s(str) {}
    bool lit(const char *t) { size_t n = strlen(t); if (s.compare(i, n, t) == 0) { i += n; return true; } return false; }
    bool ch(char c) { if (i < s.size() && s[i] == c) { ++i; return true; } return false; }
    char peek(size_t o = 0) const { return i + o < s.size() ? s[i + o] : '\0'; }
    bool digits(long &v) {
        size_t j = i; long r = 0;
        while (j < s.size() && isdigit((unsigned char)s[j])) { r = r * 10 + (s[j] - '0'); if (r > 100000000L) r = 100000000L; ++j; }
        if (j == i) return false; v = r; i = j; return true;
    }
    // {\w+}
    bool name(std::string &out) {
        if (peek() != '{') return false;
        size_t j = i + 1;
        while (j < s.size() && (isalnum((unsigned char)s[j]) || s[j] == '_')) ++j;
        if (j == i + 1 || j >= s.size() || s[j] != '}') return false;
        out = s.substr(i + 1, j - i - 1); i = j + 1; return true;
    }
    // (\d+(\.\d*)?|\.\d+)
    bool number(double &v) {
        size_t j = i;
        if (j < s.size() && isdigit((unsigned char)s[j])) {
            while (j < s.size() && isdigit((unsigned char)s[j])) ++j;
            if (j < s.size() && s[j] == '.') { ++j; while (j < s.size() && isdigit((unsigned char)s[j])) ++j; }
        } else if (j < s.size() && s[j] == '.' && j + 1 < s.size() && isdigit((unsigned char)s[j + 1])) {
            ++j; while (j < s.size() && isdigit((unsigned char)s[j])) ++j;
        } else return false;
        v = atof(s.substr(i, j - i).c_str()); i = j; return true;
    }
    // ,(\d+),(\d+)  - all or nothing
    bool comma_pair(long &a, long &b) {
        size_t save = i;
        if (ch(',') && digits(a) && ch(',') && digits(b)) return true;
        i = save; return false;
    }
};

static inline const int *dim_map() { static const int m[4] = {0, 2, 3, 1}; return m; }   // VGSL (b,h,w,c) -> NCHW

// model.py:53-64
static inline void name_block(const std::string &tok, const std::string &layer, const std::string &given, bool has_given,
                              int idx, std::string &name, std::string &block) {
    if (has_given) name = given;
    else {
        std::string l;
        bool in_nonword = false;
        for (char c : layer) {
            if (isalnum((unsigned char)c) || c == '_') { l.push_back(c); in_nonword = false; }
            else if (!in_nonword) { l.push_back('_'); in_nonword = true; }
        }
        name = l + "_" + std::to_string(idx);
    }
    std::string stripped = tok;
    size_t a = stripped.find('{'), b = stripped.rfind('}');
    if (a != std::string::npos && b != std::string::npos && b > a + 1) stripped.erase(a, b - a + 1);   // \{.+\} greedy
    size_t p = 0;
    while (p < stripped.size() && !isdigit((unsigned char)stripped[p])) ++p;
    if (p == 0) block = "{" + name + "}" + stripped;            // no leading non-digit run: re.split yields [block]
    else block = stripped.substr(0, p) + "{" + name + "}" + stripped.substr(p);
}

static inline int conv_static(int n, int k, int s, int d, int p) {
    if (n == 0) return 0;
    double v = std::floor((double)(n + 2 * p - d * (k - 1) - 1) / s + 1);
    return (int)std::max(v, 1.0);
}
static inline int pool_static(int n, int k, int s) {
    if (n == 0) return 0;
    return (int)std::floor((double)(n - (k - 1) - 1) / s + 1);
}

// layers.py:313-333 on shapes only.  perm5 (optional) receives the 5-D permutation, dest the merge position.
static inline void reshape_dims(const int64_t in[4], const Node &n, int64_t out[4], int perm5[5], int *dest_out, int64_t shape5[5]) {
    int src = n.rs_src;
    int64_t a = n.rs_a, b = n.rs_b;
    if (a == -1) { if (b <= 0 || in[src] % b) throw ShapeError("reshape: dimension not divisible"); a = in[src] / b; }
    else if (b == -1) { if (a <= 0 || in[src] % a) throw ShapeError("reshape: dimension not divisible"); b = in[src] / a; }
    if (a * b != in[src]) throw ShapeError("reshape: invalid split " + std::to_string(a) + "x" + std::to_string(b) +
                                           " of dimension of size " + std::to_string(in[src]));
    int64_t s5[5]; int j = 0;
    for (int i = 0; i < 4; ++i) { if (i == src) { s5[j++] = a; s5[j++] = b; } else s5[j++] = in[i]; }
    int dest = n.rs_low, s = src;
    if (n.rs_high != src) dest = n.rs_high; else s += 1;
    int perm[5] = {0, 1, 2, 3, 4};
    int step = dest > s ? 1 : -1;
    for (int x = s; x != dest; x += step) std::swap(perm[x], perm[x + step]);
    int64_t p5[5];
    for (int i = 0; i < 5; ++i) p5[i] = s5[perm[i]];
    if (dest < 0 || dest > 3) throw ShapeError("reshape: invalid destination");
    j = 0;
    for (int i = 0; i < 5; ++i) {
        if (i == dest) { out[j++] = p5[i] * p5[i + 1]; ++i; }
        else out[j++] = p5[i];
    }
    if (perm5) for (int i = 0; i < 5; ++i) perm5[i] = perm[i];
    if (dest_out) *dest_out = dest;
    if (shape5) for (int i = 0; i < 5; ++i) shape5[i] = s5[i];
}

// ---------------------------------------------------------------------------------------------
// parser
// ---------------------------------------------------------------------------------------------
class Parser {
public:
    int idx = -1;

    std::unique_ptr<Node> leaf(const int shape[4], const std::string &tok) {
        std::unique_ptr<Node> n;
        if ((n = addition(shape, tok))) return n;
        if ((n = identity(shape, tok))) return n;
        if ((n = rnn(shape, tok))) return n;
        if ((n = dropout(shape, tok))) return n;
        if ((n = maxpool(shape, tok))) return n;
        if ((n = conv(shape, tok))) return n;
        if ((n = output(shape, tok))) return n;
        if ((n = reshape(shape, tok))) return n;
        wav2vec(tok);
        if ((n = groupnorm(shape, tok))) return n;
        return nullptr;
    }

    std::unique_ptr<Node> sequence(const int shape[4], const std::vector<std::string> &toks, bool parallel) {
        auto node = std::make_unique<Node>();
        node->kind = parallel ? K_PARALLEL : K_SERIES;
        memcpy(node->in_shape, shape, sizeof(int) * 4);
        int cur[4]; memcpy(cur, shape, sizeof(cur));
        int prev[4] = {0, 0, 0, 0}; bool have_prev = false; int channels = 0;
        size_t i = 0;
        while (i < toks.size()) {
            const std::string &tok = toks[i];
            size_t span = 1;
            std::unique_ptr<Node> child = leaf(cur, tok);
            if (!child && !tok.empty() && tok[0] == '[') child = group(cur, toks, i, false, span);
            else if (!child && !tok.empty() && tok[0] == '(') child = group(cur, toks, i, true, span);
            if (!child) throw SpecError(tok + " invalid layer definition");
            if (parallel) {
                if (have_prev && (prev[2] != child->out_shape[2] || prev[3] != child->out_shape[3]))
                    throw SpecError("Output shape in parallel block not equal!");
                memcpy(prev, child->out_shape, sizeof(prev)); have_prev = true;
                channels += child->out_shape[1];
            } else memcpy(cur, child->out_shape, sizeof(cur));
            node->children.push_back(std::move(child));
            i += span;
        }
        if (parallel) {
            if (!have_prev) throw SpecError("empty parallel block");
            node->out_shape[0] = prev[0]; node->out_shape[1] = channels; node->out_shape[2] = prev[2]; node->out_shape[3] = prev[3];
        } else memcpy(node->out_shape, cur, sizeof(cur));
        std::vector<Node *> lv; leaves(node.get(), lv);
        for (size_t k = 0; k < lv.size(); ++k) { if (k) node->name += " "; node->name += lv[k]->name; }
        return node;
    }

    static void leaves(Node *n, std::vector<Node *> &out) {
        if (n->kind == K_SERIES || n->kind == K_PARALLEL) { for (auto &c : n->children) leaves(c.get(), out); }
        else out.push_back(n);
    }

private:
    static int depth(const std::string &t, char open_c, char close_c, char other_open, char other_close) {
        int r = 0;
        for (size_t i = 0; i < t.size(); ++i) { if (t[i] == open_c) ++r; else if (t[i] != other_open) break; }
        for (size_t i = t.size(); i-- > 0;) { if (t[i] == close_c) --r; else if (t[i] != other_close) break; }
        return r;
    }

    std::unique_ptr<Node> group(const int shape[4], const std::vector<std::string> &toks, size_t i, bool parallel, size_t &span) {
        char oc = parallel ? '(' : '[', cc = parallel ? ')' : ']';
        char xo = parallel ? '[' : '(', xc = parallel ? ']' : ')';
        std::vector<std::string> inner;
        const std::string &first = toks[i];
        if (first.back() == cc) { inner.push_back(first.substr(1, first.size() - 2)); span = 1; }
        else {
            int d = 0; size_t j = 0; bool closed = false;
            for (j = 0; i + j < toks.size(); ++j) { d += depth(toks[i + j], oc, cc, xo, xc); if (d == 0) { closed = true; break; } }
            if (!closed || d) throw SpecError("Unbalanced parentheses in VGSL spec");
            inner.push_back(first.substr(1));
            for (size_t k = 1; k < j; ++k) inner.push_back(toks[i + k]);
            const std::string &last = toks[i + j];
            inner.push_back(last.substr(0, last.size() - 1));
            span = j + 1;
        }
        auto node = sequence(shape, inner, parallel);
        std::vector<Node *> lv; leaves(node.get(), lv);
        if (lv.empty()) throw SpecError("empty block in VGSL spec");
        lv.front()->block = std::string(1, oc) + lv.front()->block;
        lv.back()->block += cc;
        return node;
    }

    std::unique_ptr<Node> mk(int kind, const int shape[4], const std::string &tok, const char *layer,
                             const std::string &nm, bool has_nm) {
        auto n = std::make_unique<Node>();
        n->kind = kind;
        memcpy(n->in_shape, shape, sizeof(int) * 4);
        memcpy(n->out_shape, shape, sizeof(int) * 4);
        ++idx;
        name_block(tok, layer, nm, has_nm, idx, n->name, n->block);
        return n;
    }

    std::unique_ptr<Node> addition(const int shape[4], const std::string &tok) {
        Cur c(tok); std::string nm; long dim, chunk;
        if (!c.ch('A')) return nullptr;
        bool has = c.name(nm);
        if (!c.digits(dim) || !c.ch(',') || !c.digits(chunk)) return nullptr;
        if (dim > 3) throw SpecError("Invalid dimension " + std::to_string(dim) + " in addition block");
        auto n = mk(K_ADD, shape, tok, "A", nm, has);
        n->add_dim = dim_map()[dim]; n->add_chunk = (int)chunk;
        n->out_shape[n->add_dim] = (int)chunk;
        return n;
    }
    std::unique_ptr<Node> identity(const int shape[4], const std::string &tok) {
        Cur c(tok); std::string nm;
        if (!c.ch('I')) return nullptr;
        bool has = c.name(nm);
        return mk(K_IDENTITY, shape, tok, "I", nm, has);
    }
    std::unique_ptr<Node> rnn(const int shape[4], const std::string &tok) {
        Cur c(tok); std::string nm; long hid;
        char type = c.peek();
        if (type != 'L' && type != 'G') return nullptr;
        c.i++;
        char dir = c.peek(); if (dir != 'f' && dir != 'r' && dir != 'b') return nullptr; c.i++;
        char dim = c.peek(); if (dim != 'x' && dim != 'y') return nullptr; c.i++;
        bool sum = c.ch('s');
        int legacy = 0;
        if (c.peek() == 'c') { legacy = 1; c.i++; } else if (c.peek() == 'o') { legacy = 2; c.i++; }
        bool has = c.name(nm);
        if (!c.digits(hid)) {
            // regex backtracking: the optional 's'/'c'/'o' cannot be re-read as anything else, so no match
            return nullptr;
        }
        auto n = mk(K_LSTM, shape, tok, type == 'L' ? "L" : "G", nm, has);
        n->hidden = (int)hid; n->bidi = dir == 'b'; n->transpose = dim == 'y'; n->summarize = sum; n->legacy = legacy;
        n->cin = shape[1];
        int osz = n->bidi ? 2 * (int)hid : (int)hid;
        n->out_shape[0] = shape[0]; n->out_shape[1] = osz;
        if (sum) { if (n->transpose) { n->out_shape[2] = 1; n->out_shape[3] = shape[3]; } else { n->out_shape[2] = shape[2]; n->out_shape[3] = 1; } }
        else { n->out_shape[2] = shape[2]; n->out_shape[3] = shape[3]; }
        return n;
    }
    std::unique_ptr<Node> dropout(const int shape[4], const std::string &tok) {
        Cur c(tok); std::string nm;
        if (!c.lit("Do")) return nullptr;
        bool has = c.name(nm);
        double p = 0.5; long dim = 1;
        c.number(p);
        { size_t save = c.i; long d; if (c.ch(',') && c.digits(d)) dim = d; else c.i = save; }
        auto n = mk(K_DROPOUT, shape, tok, "Do", nm, has);
        n->drop_p = p; n->drop_dim = (int)dim;
        return n;
    }
    std::unique_ptr<Node> maxpool(const int shape[4], const std::string &tok) {
        Cur c(tok); std::string nm; long ky, kx, sy, sx;
        if (!c.lit("Mp")) return nullptr;
        bool has = c.name(nm);
        if (!c.digits(ky) || !c.ch(',') || !c.digits(kx)) return nullptr;
        bool st = c.comma_pair(sy, sx);
        auto n = mk(K_POOL, shape, tok, "Mp", nm, has);
        n->kh = (int)ky; n->kw = (int)kx;
        // reference: `kernel if not m.group(5) else int(m.group(5))` - a literal stride of "0" is a truthy string
        n->sy = st ? (int)sy : (int)ky; n->sx = st ? (int)sx : (int)kx;
        if (n->sy <= 0 || n->sx <= 0 || n->kh <= 0 || n->kw <= 0) throw SpecError(tok + ": pooling kernel/stride must be positive");
        n->out_shape[2] = pool_static(shape[2], n->kh, n->sy);
        n->out_shape[3] = pool_static(shape[3], n->kw, n->sx);
        return n;
    }
    std::unique_ptr<Node> conv(const int shape[4], const std::string &tok) {
        Cur c(tok); std::string nm; long ky, kx, out, sy = 1, sx = 1, dy = 1, dx = 1;
        if (!c.ch('C')) return nullptr;
        bool transposed = c.ch('T');
        int act;
        char a = c.peek();
        if (a == 's') { act = ACT_SIGMOID_LOGITS; c.i++; }
        else if (a == 't') { act = ACT_TANH; c.i++; }
        else if (a == 'r') { act = ACT_RELU; c.i++; }
        else if (a == 'l') { c.i++; if (c.peek() == 'r') { act = ACT_LEAKY; c.i++; } else act = ACT_LINEAR; }
        else if (a == 'm') { act = ACT_SOFTMAX; c.i++; }
        else return nullptr;
        bool has = c.name(nm);
        if (!c.digits(ky) || !c.ch(',') || !c.digits(kx) || !c.ch(',') || !c.digits(out)) return nullptr;
        bool has_stride = c.comma_pair(sy, sx);
        bool has_dil = has_stride ? c.comma_pair(dy, dx) : false;
        if (!has_stride) { sy = sx = 1; }
        if (!has_dil) { dy = dx = 1; }
        if (transposed) throw Unsupported(tok + ": transposed convolution is not on the rpred/blla path and is not implemented by the engine");
        if (ky <= 0 || kx <= 0 || out <= 0 || sy <= 0 || sx <= 0 || dy <= 0 || dx <= 0) throw SpecError(tok + ": convolution parameters must be positive");
        auto n = mk(K_CONV, shape, tok, "C", nm, has);
        n->kh = (int)ky; n->kw = (int)kx; n->cout = (int)out; n->cin = shape[1];
        n->sy = (int)sy; n->sx = (int)sx; n->dy = (int)dy; n->dx = (int)dx;
        n->py = (n->dy * (n->kh - 1)) / 2; n->px = (n->dx * (n->kw - 1)) / 2; n->act = act;
        n->out_shape[1] = n->cout;
        n->out_shape[2] = conv_static(shape[2], n->kh, n->sy, n->dy, n->py);
        n->out_shape[3] = conv_static(shape[3], n->kw, n->sx, n->dx, n->px);
        return n;
    }
    std::unique_ptr<Node> output(const int shape[4], const std::string &tok) {
        Cur c(tok); std::string nm; long out;
        if (!c.ch('O')) return nullptr;
        bool has = c.name(nm);
        char d = c.peek(); if (d != '2' && d != '1' && d != '0') return nullptr; c.i++;
        char t = c.peek(); if (t != 'l' && t != 's' && t != 'c') return nullptr; c.i++;
        bool aug = c.ch('a');
        if (!c.digits(out)) return nullptr;
        int dim = d - '0';
        if (dim == 0) throw SpecError("categorical output not supported, yet.");
        if (t == 'c' && dim == 2) throw SpecError("CTC not supported for heatmap output");
        if (!(((t == 'l' || t == 's') && out >= 1) || t == 'c')) throw SpecError("unsupported output specification");
        if (dim == 2) {
            auto n = mk(K_CONV, shape, tok, std::string(1, t).c_str(), nm, has);
            n->kh = n->kw = 1; n->sy = n->sx = n->dy = n->dx = 1; n->py = n->px = 0;
            n->cin = shape[1]; n->cout = (int)out; n->act = t == 'l' ? ACT_SIGMOID_LOGITS : ACT_SOFTMAX;
            n->out_shape[1] = (int)out;
            return n;
        }
        auto n = mk(K_LINEAR, shape, tok, "O", nm, has);
        n->cin = shape[1]; n->cout = (int)out; n->aug = aug;
        n->out_shape[1] = (int)out;
        return n;
    }
    std::unique_ptr<Node> reshape(const int shape[4], const std::string &tok) {
        Cur c(tok); std::string nm; long d, a, b, hi, lo;
        if (!c.ch('S')) return nullptr;
        bool has = c.name(nm);
        if (!c.digits(d) || !c.ch('(') || !c.digits(a) || !c.ch('x') || !c.digits(b) || !c.ch(')') ||
            !c.digits(hi) || !c.ch(',') || !c.digits(lo)) return nullptr;
        if (d > 3 || hi > 3 || lo > 3) throw SpecError(tok + ": invalid dimension in reshape block");   // reference: KeyError
        if (a == 0) a = -1; else if (b == 0) b = -1;
        if (d != hi && d != lo)
            throw SpecError("Either high (" + std::to_string(hi) + ") or low (" + std::to_string(lo) + ") must be source dimension (" + std::to_string(d) + ")");
        if (a == -1 && b == -1) throw SpecError("Only one size may be -1");
        auto n = mk(K_RESHAPE, shape, tok, "S", nm, has);
        n->rs_src = dim_map()[d]; n->rs_a = (int)a; n->rs_b = (int)b; n->rs_high = dim_map()[hi]; n->rs_low = dim_map()[lo];
        int64_t in[4], out[4];
        for (int i = 0; i < 4; ++i) in[i] = shape[i] ? shape[i] : 1;          // probe with variable dims = 1
        try { reshape_dims(in, *n, out, nullptr, nullptr, nullptr); }
        catch (ShapeError &e) { throw SpecError(tok + ": " + e.what()); }
        for (int i = 0; i < 4; ++i) n->out_shape[i] = (int)out[i];
        return n;
    }
    void wav2vec(const std::string &tok) {
        Cur c(tok); std::string nm; long a, b, d; double p;
        if (!c.ch('W')) return;
        if (!c.name(nm)) return;
        if (c.digits(a) && c.ch(',') && c.digits(b) && c.ch(',') && c.number(p) && c.ch(',') && c.digits(d))
            throw Unsupported(tok + ": wav2vec2 masking layers are pre-training only and not implemented by the engine");
    }
    std::unique_ptr<Node> groupnorm(const int shape[4], const std::string &tok) {
        Cur c(tok); std::string nm; long g;
        if (!c.lit("Gn")) return nullptr;
        bool has = c.name(nm);
        if (!c.digits(g)) return nullptr;
        if (g <= 0 || shape[1] % g) throw SpecError(tok + ": num_channels must be divisible by num_groups");
        auto n = mk(K_GN, shape, tok, "Gn", nm, has);
        n->groups = (int)g; n->cin = shape[1];
        return n;
    }
};

// ---------------------------------------------------------------------------------------------
// Plan = parsed model
// ---------------------------------------------------------------------------------------------
struct Plan {
    std::string spec, named_spec;
    int input[4] = {0, 0, 0, 0};      // (batch, channels, height, width)
    int output[4] = {0, 0, 0, 0};
    std::unique_ptr<Node> root;
    std::vector<Node *> leaf_nodes;
    std::vector<TensorDecl> tensors;
};

static inline void assign_paths(Node *n, const std::string &prefix) {
    for (auto &c : n->children) {
        c->path = prefix + c->name;
        if (c->kind == K_SERIES || c->kind == K_PARALLEL) assign_paths(c.get(), c->path + ".");
    }
}

static inline std::unique_ptr<Plan> parse_spec(const std::string &spec_in) {
    auto plan = std::make_unique<Plan>();
    std::string spec = spec_in;
    size_t b = spec.find_first_not_of(" \t\r\n"), e = spec.find_last_not_of(" \t\r\n");
    if (b == std::string::npos) throw SpecError("vgsl specification argument is missing in args.");
    spec = spec.substr(b, e - b + 1);
    plan->spec = spec;
    if (spec.front() != '[' || spec.back() != ']') throw SpecError("Non-sequential models not supported");
    std::string body = spec.substr(1, spec.size() - 2);
    std::vector<std::string> toks;
    { size_t p = 0; while (true) { size_t q = body.find(' ', p); toks.push_back(body.substr(p, q == std::string::npos ? q : q - p)); if (q == std::string::npos) break; p = q + 1; } }
    Cur c(toks[0]); long bb, hh, ww, cc;
    if (!(c.digits(bb) && c.ch(',') && c.digits(hh) && c.ch(',') && c.digits(ww) && c.ch(',') && c.digits(cc)))
        throw SpecError("Invalid input spec.");
    plan->input[0] = (int)bb; plan->input[1] = (int)cc; plan->input[2] = (int)hh; plan->input[3] = (int)ww;
    std::vector<std::string> rest(toks.begin() + 1, toks.end());
    Parser p;
    plan->root = p.sequence(plan->input, rest, false);
    assign_paths(plan->root.get(), "");
    memcpy(plan->output, plan->root->out_shape, sizeof(int) * 4);
    Parser::leaves(plan->root.get(), plan->leaf_nodes);
    plan->named_spec = "[" + toks[0];
    for (size_t i = 0; i < plan->leaf_nodes.size(); ++i) {
        Node *n = plan->leaf_nodes[i];
        n->leaf_index = (int)i;
        plan->named_spec += " " + n->block;
        std::string pre = "nn." + n->path;
        auto add = [&](const std::string &sfx, std::vector<int64_t> shp, int slot) { plan->tensors.push_back({pre + sfx, std::move(shp), (int)i, slot}); };
        if (n->kind == K_CONV) { add(".co.weight", {n->cout, n->cin, n->kh, n->kw}, 0); add(".co.bias", {n->cout}, 1); }
        else if (n->kind == K_LINEAR) { add(".lin.weight", {n->cout, n->cin + (n->aug ? 1 : 0)}, 0); add(".lin.bias", {n->cout}, 1); }
        else if (n->kind == K_GN) { add(".layer.weight", {n->cin}, 0); add(".layer.bias", {n->cin}, 1); }
        else if (n->kind == K_LSTM && n->legacy) {
            // legacy cells (layers.py:498-511): a constant-one input column instead of biases; 'c' = nn.LSTM(in + 1, h, bias=False),
            // 'o' = PeepholeBidiLSTM (always both directions; peephole vectors weight_{ip,fp,op})
            int64_t h4 = 4 * (int64_t)n->hidden;
            const int ndir = (n->legacy == 2 || n->bidi) ? 2 : 1;
            for (int d = 0; d < ndir; ++d) {
                const std::string sfx = d ? "_reverse" : "";
                add(".layer.weight_ih_l0" + sfx, {h4, n->cin + 1}, d * 5 + 0); add(".layer.weight_hh_l0" + sfx, {h4, n->hidden}, d * 5 + 1);
                if (n->legacy == 2) {
                    add(".layer.weight_ip_l0" + sfx, {n->hidden}, d * 5 + 2); add(".layer.weight_fp_l0" + sfx, {n->hidden}, d * 5 + 3);
                    add(".layer.weight_op_l0" + sfx, {n->hidden}, d * 5 + 4);
                }
            }
        }
        else if (n->kind == K_LSTM && !n->legacy) {
            int64_t h4 = 4 * (int64_t)n->hidden;
            // nn.LSTM parameter order: all forward tensors, then the reverse ones
            add(".layer.weight_ih_l0", {h4, n->cin}, 0); add(".layer.weight_hh_l0", {h4, n->hidden}, 1);
            add(".layer.bias_ih_l0", {h4}, 2); add(".layer.bias_hh_l0", {h4}, 3);
            if (n->bidi) {
                add(".layer.weight_ih_l0_reverse", {h4, n->cin}, 4); add(".layer.weight_hh_l0_reverse", {h4, n->hidden}, 5);
                add(".layer.bias_ih_l0_reverse", {h4}, 6); add(".layer.bias_hh_l0_reverse", {h4}, 7);
            }
        }
    }
    plan->named_spec += "]";
    return plan;
}

// ---------------------------------------------------------------------------------------------
// runtime dims + seq_len arithmetic (host integers; floats where the reference uses float tensors)
// ---------------------------------------------------------------------------------------------
struct Dims { int64_t n = 0, c = 0, h = 0, w = 0; };

static inline int64_t conv_runtime(int64_t in, int k, int s, int d, int p) {
    int64_t v = in + 2 * p - (int64_t)d * (k - 1) - 1;
    if (v < 0) throw ShapeError("convolution: kernel larger than padded input");
    return v / s + 1;
}

// seq_len' of one leaf for one line (the reference computes these on int/float32 tensors)
static inline int32_t leaf_len(const Node &n, int32_t L, const Dims &in, const Dims &out) {
    switch (n.kind) {
    case K_CONV: {   // clamp(floor((L + 2p - d(k-1) - 1).float() / s + 1), min=1).int()   layers.py:858-859
        float v = std::floor((float)(L + 2 * n.px - n.dx * (n.kw - 1) - 1) / (float)n.sx + 1.0f);
        if (v < 1.0f) v = 1.0f;
        return (int32_t)v;
    }
    case K_POOL: {   // floor((L - (k-1) - 1).float() / s + 1).int()                       layers.py:387
        float v = std::floor((float)(L - (n.kw - 1) - 1) / (float)n.sx + 1.0f);
        return (int32_t)v;
    }
    case K_RESHAPE: { // (seq_len * (float(initial_len) / o.shape[3])).int()               layers.py:334
        float ratio = (float)((double)in.w / (double)out.w);
        return (int32_t)((float)L * ratio);
    }
    default: return L;
    }
}

static inline Dims leaf_dims(const Node &n, const Dims &in) {
    Dims o = in;
    switch (n.kind) {
    case K_CONV:
        if (in.c != n.cin) throw ShapeError(n.name + ": expected " + std::to_string(n.cin) + " input channels, got " + std::to_string(in.c));
        o.c = n.cout; o.h = conv_runtime(in.h, n.kh, n.sy, n.dy, n.py); o.w = conv_runtime(in.w, n.kw, n.sx, n.dx, n.px); break;
    case K_POOL:
        if (in.h < n.kh || in.w < n.kw) throw ShapeError(n.name + ": input smaller than pooling window");
        o.h = (in.h - n.kh) / n.sy + 1; o.w = (in.w - n.kw) / n.sx + 1; break;
    case K_RESHAPE: {
        int64_t i4[4] = {in.n, in.c, in.h, in.w}, o4[4];
        reshape_dims(i4, n, o4, nullptr, nullptr, nullptr);
        o.n = o4[0]; o.c = o4[1]; o.h = o4[2]; o.w = o4[3]; break;
    }
    case K_LSTM:
        if (in.c != n.cin) throw ShapeError(n.name + ": expected " + std::to_string(n.cin) + " input features, got " + std::to_string(in.c));
        o.c = n.bidi ? 2 * n.hidden : n.hidden;
        if (n.summarize) { if (n.transpose) o.h = 1; else o.w = 1; }
        break;
    case K_LINEAR:
        if (in.c != n.cin) throw ShapeError(n.name + ": expected " + std::to_string(n.cin) + " input features, got " + std::to_string(in.c));
        o.c = n.cout; break;
    case K_GN:
        if (in.c != n.cin) throw ShapeError(n.name + ": expected " + std::to_string(n.cin) + " channels, got " + std::to_string(in.c));
        break;
    case K_ADD: {
        int64_t *d[4] = {&o.n, &o.c, &o.h, &o.w};
        if (*d[n.add_dim] < n.add_chunk) throw ShapeError(n.name + ": addition chunk larger than dimension");
        *d[n.add_dim] = n.add_chunk; break;
    }
    default: break;
    }
    return o;
}

}  // namespace kb
