// line_prep.cuh - bbox line extraction and the PIL half of the reference's input transforms on the device.
//
//   im.crop(box)                                   kraken/lib/segmentation.py:1631-1643 (bbox lines, horizontal text)
//   v2.Grayscale / pil_to_mode('L')                kraken/lib/dataset/utils.py:123-129
//   pil_fixed_resize: img.resize((int(w*oh/h), oh), Resampling.LANCZOS)      kraken/lib/functional_im_transforms.py:58-82
//   v2.Pad(pad, fill=255)                          kraken/lib/dataset/utils.py:146-147
//
// The arithmetic of these steps is Pillow's (third-party, `pillow>=9.2.0` in the reference's pyproject.toml): Convert.c rgb2l and
// Resample.c for 8-bit images.  Restated (oracle/pil_resample.py has the same restatement in numpy, pinned against Pillow itself):
//   L = (R * 19595 + G * 38470 + B * 7471 + 0x8000) >> 16
//   per axis and output pixel xx: scale = in / out, filterscale = max(scale, 1), support = 3 filterscale, center = (xx + 0.5) scale,
//   window [xmin, xmin + n) = [max(int(center - support + 0.5), 0), min(int(center + support + 0.5), in)),
//   w[x] = lanczos((x + xmin - center + 0.5) / filterscale) / sum(w)  in double,  k[x] = (int)(+-0.5 + w[x] * 2^22),
//   out = clip8((2^21 + sum_x in[xmin + x] * k[x]) >> 22);  horizontal pass over the rows the vertical pass reads, then vertical pass.
// Split of the work: the coefficient windows (double-precision sinc through the host's libm - the very function Pillow calls, so the
// fixed-point integers are identical by construction) are computed on the host, one std::thread per slice of lines; the integer
// convolutions - all of the byte traffic - run on the device.  A pass Pillow skips (size unchanged) runs here with the identity window
// k = {2^22}, which reproduces the input exactly.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <thread>
#include <vector>

namespace kb {
namespace lp {

constexpr int PRECISION_BITS = 32 - 8 - 2;

// per line, in int32 words (device copy of the host plan)
struct LineMeta {
    int x0, y0, in_w, in_h;          // crop box
    int ow, width;                   // resized width, padded width
    int hb_off, hk_off, ksize_h;     // horizontal bounds [ow][2] / coefficients [ow][ksize_h] (word offsets into the table)
    int vb_off, vk_off, ksize_v;     // vertical   bounds [out_h][2] / coefficients [out_h][ksize_v]
    int row_first, rows;             // source rows the vertical pass reads: [row_first, row_first + rows)
    long long tmp_off;               // byte offset of the line's horizontal-pass output [rows][ow]
};

inline double sinc_filter(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return sin(x) / x;
}
inline double lanczos_filter(double x) {
    if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
    return 0.0;
}
inline int axis_ksize(int in_size, int out_size) {
    double filterscale = (double)in_size / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    return (int)ceil(3.0 * filterscale) * 2 + 1;
}
// Resample.c precompute_coeffs + normalize_coeffs_8bpc for box = the whole axis.  bounds: [out][2] = (xmin, count); kk: [out][ksize]
inline void axis_coeffs(int in_size, int out_size, int ksize, int32_t *bounds, int32_t *kk) {
    if (in_size == out_size) {                       // Pillow skips the pass: identity window
        for (int xx = 0; xx < out_size; ++xx) {
            bounds[2 * xx] = xx; bounds[2 * xx + 1] = 1;
            kk[(size_t)xx * ksize] = 1 << PRECISION_BITS;
            for (int x = 1; x < ksize; ++x) kk[(size_t)xx * ksize + x] = 0;
        }
        return;
    }
    const double scale = (double)in_size / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 3.0 * filterscale, ss = 1.0 / filterscale;
    std::vector<double> k((size_t)ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            const double w = lanczos_filter((x + xmin - center + 0.5) * ss);
            k[x] = w; ww += w;
        }
        for (int x = 0; x < xmax; ++x) if (ww != 0.0) k[x] /= ww;
        int32_t *ko = kk + (size_t)xx * ksize;
        for (int x = 0; x < xmax; ++x) ko[x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << PRECISION_BITS)) : (int)(0.5 + k[x] * (1 << PRECISION_BITS));
        for (int x = xmax; x < ksize; ++x) ko[x] = 0;
        bounds[2 * xx] = xmin; bounds[2 * xx + 1] = xmax;
    }
}
inline int identity_ksize() { return 1; }

__device__ __forceinline__ int clip8_fixed(int v) { v >>= PRECISION_BITS; return v < 0 ? 0 : (v > 255 ? 255 : v); }

// horizontal pass fused with the crop and the grayscale conversion: tmp[line][r][xx], r over the rows the vertical pass needs
__global__ void k_prep_horizontal(const uint8_t *__restrict__ page, int page_w, int channels, const LineMeta *__restrict__ meta,
                                  const int32_t *__restrict__ tab, uint8_t *__restrict__ tmp) {
    const LineMeta L = meta[blockIdx.z];
    const int r = blockIdx.y, xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= L.rows || xx >= L.ow) return;
    const int xmin = tab[L.hb_off + 2 * xx], cnt = tab[L.hb_off + 2 * xx + 1];
    const int32_t *k = tab + L.hk_off + (size_t)xx * L.ksize_h;
    const uint8_t *src = page + ((size_t)(L.y0 + L.row_first + r) * page_w + (size_t)(L.x0 + xmin)) * channels;
    int ss = 1 << (PRECISION_BITS - 1);
    if (channels == 1) {
        for (int x = 0; x < cnt; ++x) ss += (int)src[x] * k[x];
    } else {
        for (int x = 0; x < cnt; ++x) {
            const unsigned l = ((unsigned)src[3 * x] * 19595u + (unsigned)src[3 * x + 1] * 38470u + (unsigned)src[3 * x + 2] * 7471u + 0x8000u) >> 16;
            ss += (int)l * k[x];
        }
    }
    tmp[L.tmp_off + (size_t)r * L.ow + xx] = (uint8_t)clip8_fixed(ss);
}

// vertical pass + white padding: lines[line][yy][c], c < width.  Also the line's maximum (tensor_invert's im.max()).
__global__ void k_prep_vertical(const uint8_t *__restrict__ tmp, const LineMeta *__restrict__ meta, const int32_t *__restrict__ tab,
                                uint8_t *__restrict__ lines, int out_h, int wmax, int pad, int *__restrict__ line_max) {
    const LineMeta L = meta[blockIdx.z];
    const int yy = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
    int v = -1;
    if (c < L.width) {
        if (c < pad || c >= pad + L.ow) v = 255;
        else {
            const int xx = c - pad;
            const int ymin = tab[L.vb_off + 2 * yy] - L.row_first, cnt = tab[L.vb_off + 2 * yy + 1];
            const int32_t *k = tab + L.vk_off + (size_t)yy * L.ksize_v;
            const uint8_t *src = tmp + L.tmp_off + (size_t)ymin * L.ow + xx;
            int ss = 1 << (PRECISION_BITS - 1);
            for (int y = 0; y < cnt; ++y) ss += (int)src[(size_t)y * L.ow] * k[y];
            v = clip8_fixed(ss);
        }
        lines[((size_t)blockIdx.z * out_h + yy) * wmax + c] = (uint8_t)v;
    }
    if (line_max) {
        for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        if ((threadIdx.x & 31) == 0 && v >= 0) atomicMax(&line_max[blockIdx.z], v);
    }
}

// `ow = int(w * oh / h)` (functional_im_transforms.py:77-78: Python int() of a true division)
inline int resized_width(int box_w, int box_h, int out_h) {
    if (box_w <= 0 || box_h <= 0 || out_h <= 0) return 0;
    return (int)((double)((long long)box_w * out_h) / (double)box_h);
}

struct Plan {
    std::vector<LineMeta> meta;
    std::vector<int32_t> tab;
    size_t tmp_bytes = 0;
    int max_rows = 0, max_ow = 0, max_width = 0;
};

// host plan of one call: per-line geometry and the coefficient tables (vertical tables shared between lines of equal box height)
inline void build_plan(Plan &pl, int n, const int32_t *boxes, int out_h, int pad) {
    pl.meta.assign((size_t)n, LineMeta());
    size_t words = 0;
    std::vector<int> vrep((size_t)n, -1);             // first line with the same box height
    for (int i = 0; i < n; ++i) {
        LineMeta &L = pl.meta[(size_t)i];
        L.x0 = boxes[4 * i]; L.y0 = boxes[4 * i + 1]; L.in_w = boxes[4 * i + 2] - boxes[4 * i]; L.in_h = boxes[4 * i + 3] - boxes[4 * i + 1];
        L.ow = resized_width(L.in_w, L.in_h, out_h); L.width = L.ow + 2 * pad;
        L.ksize_h = L.in_w == L.ow ? identity_ksize() : axis_ksize(L.in_w, L.ow);
        L.hb_off = (int)words; words += (size_t)2 * L.ow;
        L.hk_off = (int)words; words += (size_t)L.ow * L.ksize_h;
        for (int j = 0; j < i; ++j) if (pl.meta[(size_t)j].in_h == L.in_h) { vrep[(size_t)i] = j; break; }
        if (vrep[(size_t)i] >= 0) {
            const LineMeta &R = pl.meta[(size_t)vrep[(size_t)i]];
            L.ksize_v = R.ksize_v; L.vb_off = R.vb_off; L.vk_off = R.vk_off;
        } else {
            L.ksize_v = L.in_h == out_h ? identity_ksize() : axis_ksize(L.in_h, out_h);
            L.vb_off = (int)words; words += (size_t)2 * out_h;
            L.vk_off = (int)words; words += (size_t)out_h * L.ksize_v;
        }
    }
    pl.tab.assign(words, 0);
    auto work = [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i) {
            LineMeta &L = pl.meta[(size_t)i];
            axis_coeffs(L.in_w, L.ow, L.ksize_h, pl.tab.data() + L.hb_off, pl.tab.data() + L.hk_off);
            if (vrep[(size_t)i] < 0) axis_coeffs(L.in_h, out_h, L.ksize_v, pl.tab.data() + L.vb_off, pl.tab.data() + L.vk_off);
        }
    };
    const int nthreads = std::max(1, std::min<int>({n / 2, 16, (int)std::thread::hardware_concurrency()}));
    if (nthreads <= 1) work(0, n);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(work, (int)((long long)n * t / nthreads), (int)((long long)n * (t + 1) / nthreads));
        for (auto &t : th) t.join();
    }
    pl.tmp_bytes = 0; pl.max_rows = pl.max_ow = pl.max_width = 0;
    for (int i = 0; i < n; ++i) {
        LineMeta &L = pl.meta[(size_t)i];
        const int32_t *vb = pl.tab.data() + L.vb_off;
        L.row_first = vb[0];
        L.rows = vb[2 * (out_h - 1)] + vb[2 * (out_h - 1) + 1] - L.row_first;       // Resample.c: ybox_first .. ybox_last
        L.tmp_off = (long long)pl.tmp_bytes;
        pl.tmp_bytes += ((size_t)L.rows * L.ow + 15) & ~(size_t)15;
        pl.max_rows = std::max(pl.max_rows, L.rows); pl.max_ow = std::max(pl.max_ow, L.ow); pl.max_width = std::max(pl.max_width, L.width);
    }
}

}  // namespace lp
}  // namespace kb
