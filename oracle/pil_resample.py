"""
oracle/pil_resample.py - CPU restatement of the line pre-processing of the reference (TEST INFRASTRUCTURE ONLY: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline may import it).

The reference scales every line image with Pillow: `ImageInputTransforms` (kraken/lib/dataset/utils.py:123-152) = Grayscale ->
`pil_fixed_resize` (kraken/lib/functional_im_transforms.py:58-82: `img.resize((ow, oh), Resampling.LANCZOS)` with
`ow = int(w * oh / h)`) -> `v2.Pad(pad, fill=255)` -> PILToTensor -> ToDtype(scale) -> tensor_invert, on the crop
`im.crop(box)` of a bbox line (kraken/lib/segmentation.py:1631-1643).

The arithmetic lives in a third-party dependency that is not under /root/reference: Pillow (kraken/pyproject.toml: `pillow>=9.2.0`;
12.2 in this image), `src/libImaging/Resample.c` and `Convert.c`.  Their published algorithm for 8-bit images, restated here:

  * RGB -> L:  L = (R * 19595 + G * 38470 + B * 7471 + 0x8000) >> 16                                      (Convert.c, rgb2l)
  * resize = horizontal pass then vertical pass over 8-bit data, each with per-output-pixel coefficient windows:
        scale = in / out; filterscale = max(scale, 1); support = 3 * filterscale (LANCZOS); ksize = ceil(support) * 2 + 1
        center = (xx + 0.5) * scale; xmin = max(int(center - support + 0.5), 0); xmax = min(int(center + support + 0.5), in)
        w[x] = lanczos((x + xmin - center + 0.5) / filterscale), normalised by their sum (double arithmetic)
        fixed point: k[x] = int(+-0.5 + w[x] * 2^22), toward zero                                           (normalize_coeffs_8bpc)
        out = clip8((2^21 + sum_x in[xmin + x] * k[x]) >> 22)                                                (arithmetic shift)
    The horizontal pass only produces the source rows the vertical pass reads.  A pass is skipped when it would not change the size.

The restatement is PINNED against Pillow itself (which IS the reference's implementation and is present in this image):
tests/test_line_prep.py compares it bit for bit on random sizes.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x: float) -> float:
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """(ksize, bounds [out][2] (xmin, count), fixed-point coefficients [out][ksize] int32) of one axis (Resample.c precompute_coeffs +
    normalize_coeffs_8bpc, box = the whole axis)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img: np.ndarray, bounds, kk, axis: int) -> np.ndarray:
    """one resampling pass over axis 1 (horizontal) or 0 (vertical) of a uint8 image"""
    a = img.astype(np.int64)
    if axis == 0:
        a = a.T
    out = np.empty((a.shape[0], bounds.shape[0]), np.uint8)
    for xx in range(bounds.shape[0]):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        s = (1 << (PRECISION_BITS - 1)) + a[:, xmin:xmin + n] @ kk[xx, :n].astype(np.int64)
        out[:, xx] = np.clip(s >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out.T if axis == 0 else out


def resize_lanczos_u8(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """`Image.fromarray(img, 'L').resize((out_w, out_h), LANCZOS)` for a uint8 (H, W) array"""
    h, w = img.shape
    cur = img
    need_h, need_v = out_w != w, out_h != h
    if need_v:
        _, vb, vk = precompute_coeffs(h, out_h)
    if need_h:
        _, hb, hk = precompute_coeffs(w, out_w)
        if need_v:
            first = int(vb[0, 0]); last = int(vb[-1, 0] + vb[-1, 1])
            cur = _pass(cur[first:last], hb, hk, 1)
            vb = vb.copy(); vb[:, 0] -= first
        else:
            cur = _pass(cur, hb, hk, 1)
    if need_v:
        cur = _pass(cur, vb, vk, 0)
    return cur.copy() if cur is img else cur


def rgb_to_l(rgb: np.ndarray) -> np.ndarray:
    """(H, W, 3) uint8 -> (H, W) uint8 as `Image.convert('L')` (ITU-R 601-2 luma, Convert.c rgb2l)"""
    r, g, b = (rgb[..., i].astype(np.uint32) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def target_width(w: int, h: int, out_h: int) -> int:
    """`ow = int(w * oh / h)` (functional_im_transforms.py:77-78; Python true division)"""
    return int(w * out_h / h)


def prepare_line(page: np.ndarray, box, out_h: int, pad: int) -> np.ndarray:
    """bbox crop -> grayscale -> LANCZOS to height `out_h` -> white padding: the uint8 (1, out_h, W') image the reference's
    PILToTensor yields for a bbox line (segmentation.py:1643, dataset/utils.py:123-147)."""
    x0, y0, x1, y1 = (int(v) for v in box)
    crop = page[y0:y1, x0:x1]
    if crop.ndim == 3:
        crop = rgb_to_l(crop)
    h, w = crop.shape
    ow = target_width(w, h, out_h)
    if ow < 1:
        raise ValueError('height and width must be > 0')
    img = resize_lanczos_u8(crop, ow, out_h) if (ow, out_h) != (w, h) else crop.copy()
    if pad:
        img = np.pad(img, ((0, 0), (pad, pad)), constant_values=255)
    return img[None]
