"""Line pre-processing: bbox crop -> Grayscale -> LANCZOS resize to the model height -> white padding.

CPU: the numpy restatement (oracle/pil_resample.py) is pinned bit for bit against Pillow, which IS the reference's implementation of
these steps (kraken/lib/functional_im_transforms.py:58-82 calls `img.resize(..., LANCZOS)`); the ABI's width formula against Python's.
GPU: `kb_prepare_lines_u8` against Pillow on random pages / boxes (gray and RGB, down- and up-scaling, with and without padding), and
page + boxes -> labels end to end against the transforms the reference applies followed by the oracle network."""
import numpy as np
import pytest
import torch
from PIL import Image

import pil_resample as pr
import kraken_b200 as kb
from kraken_b200 import lineprep


def _pil_line(page: np.ndarray, box, out_h: int, pad: int) -> np.ndarray:
    """the reference's own steps with Pillow: im.crop(box) -> Grayscale -> pil_fixed_resize -> Pad(fill=255) -> PILToTensor"""
    im = Image.fromarray(page, 'L' if page.ndim == 2 else 'RGB').crop(tuple(int(v) for v in box)).convert('L')
    w, h = im.size
    ow = int(w * out_h / h)
    im = im.resize((ow, out_h), Image.LANCZOS)
    a = np.asarray(im, np.uint8)
    if pad:
        a = np.pad(a, ((0, 0), (pad, pad)), constant_values=255)
    return a[None]


def _random_boxes(rng, ph, pw, n, hmin=8, hmax=180, wmin=10):
    out = []
    while len(out) < n:
        h = int(rng.integers(hmin, min(hmax, ph) + 1)); w = int(rng.integers(wmin, pw + 1))
        y0 = int(rng.integers(0, ph - h + 1)); x0 = int(rng.integers(0, pw - w + 1))
        out.append((x0, y0, x0 + w, y0 + h))
    return out


@pytest.mark.parametrize('out_h', [48, 120])
def test_oracle_resample_is_pillow_bit_for_bit(out_h):
    rng = np.random.default_rng(out_h)
    for _ in range(25):
        h = int(rng.integers(4, 260)); w = int(rng.integers(4, 700))
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ow = pr.target_width(w, h, out_h)
        if ow < 1:
            continue
        ref = np.asarray(Image.fromarray(a, 'L').resize((ow, out_h), Image.LANCZOS))
        assert np.array_equal(pr.resize_lanczos_u8(a, ow, out_h), ref), (h, w, ow)
    # equal height: Pillow skips the vertical pass; equal size: a copy
    a = rng.integers(0, 256, (out_h, 333), dtype=np.uint8)
    assert np.array_equal(pr.resize_lanczos_u8(a, 333, out_h), a)
    assert np.array_equal(pr.resize_lanczos_u8(a, 200, out_h), np.asarray(Image.fromarray(a, 'L').resize((200, out_h), Image.LANCZOS)))


def test_oracle_prepare_line_matches_the_reference_steps():
    rng = np.random.default_rng(3)
    gray = rng.integers(0, 256, (300, 500), dtype=np.uint8)
    rgb = rng.integers(0, 256, (300, 500, 3), dtype=np.uint8)
    assert np.array_equal(pr.rgb_to_l(rgb), np.asarray(Image.fromarray(rgb, 'RGB').convert('L')))
    for page in (gray, rgb):
        for box in _random_boxes(rng, 300, 500, 8):
            for pad in (0, 16):
                assert np.array_equal(pr.prepare_line(page, box, 48, pad), _pil_line(page, box, 48, pad)), (box, pad)


def test_abi_line_width_is_the_python_formula():
    rng = np.random.default_rng(5)
    for _ in range(2000):
        w, h, oh, pad = int(rng.integers(1, 5000)), int(rng.integers(1, 400)), int(rng.choice([30, 48, 64, 120])), int(rng.choice([0, 16]))
        ow = int(w * oh / h)
        assert lineprep.line_width(w, h, oh, pad) == (ow + 2 * pad if ow >= 1 else 0), (w, h, oh)
    assert lineprep.line_width(0, 10, 48, 16) == 0 and lineprep.line_width(10, 0, 48, 16) == 0


def test_engine_host_tables_equal_the_oracles():
    """the coefficient windows the engine computes on the host (csrc/line_prep.cuh) are the oracle's, i.e. Pillow's, integer for integer"""
    import ctypes as C
    from kraken_b200._lib import check, lib
    rng = np.random.default_rng(9)
    cases = [(int(rng.integers(2, 3000)), int(rng.integers(1, 1500))) for _ in range(40)] + [(96, 48), (48, 96), (48, 48), (7, 120), (2000, 3)]
    for insz, outsz in cases:
        ks = C.c_int32()
        check(lib.kb_debug_axis_coeffs(insz, outsz, C.byref(ks), None, None, 0))
        b = np.zeros((outsz, 2), np.int32); kk = np.zeros((outsz, ks.value), np.int32)
        check(lib.kb_debug_axis_coeffs(insz, outsz, C.byref(ks), b.ctypes.data, kk.ctypes.data, kk.size))
        if insz == outsz:
            assert ks.value == 1 and np.array_equal(b[:, 0], np.arange(outsz)) and (kk == 1 << 22).all()
            continue
        oks, ob, okk = pr.precompute_coeffs(insz, outsz)
        assert ks.value == oks and np.array_equal(b, ob) and np.array_equal(kk, okk), (insz, outsz)


CFG2 = '[1,48,0,1 Cr3,3,32 Mp2,2 Cr3,3,64 Mp2,2 S1(1x0)1,3 Lbx256 O1c200]'


@pytest.mark.gpu
@pytest.mark.parametrize('mode,out_h,pad', [('L', 48, 16), ('RGB', 48, 16), ('L', 120, 0), ('RGB', 30, 0)])
def test_gpu_prepared_lines_equal_pillow(mode, out_h, pad):
    rng = np.random.default_rng(out_h + pad)
    ph, pw = 700, 1100
    page = rng.integers(0, 256, (ph, pw) if mode == 'L' else (ph, pw, 3), dtype=np.uint8)
    page[100:200, 300:800] = 255                                      # some saturated / flat areas
    boxes = _random_boxes(rng, ph, pw, 37, hmin=max(out_h // 6, 4), hmax=260)
    boxes += [(0, 0, pw, ph), (10, 20, 10 + 333, 20 + out_h), (5, 5, 5 + 64, 5 + 2 * out_h), (7, 9, 7 + 90, 9 + out_h // 2)]   # whole page, equal height, exact 1/2, exact 2x
    m = kb.TorchVGSLModel(vgsl=f'[1,{out_h},0,1 Cr3,3,8 O1c10]' if out_h != 48 else CFG2)
    m.init_weights()
    m.to('cuda:0')
    for src in (page, torch.as_tensor(page).cuda()):                  # host page (copied inside the call) and device-resident page
        lines, widths, inv = lineprep.prepare_box_lines(m, src, boxes, pad=pad)
        got = lines.cpu().numpy()
        for i, box in enumerate(boxes):
            ref = _pil_line(page, box, out_h, pad)
            assert widths[i] == ref.shape[2], (i, box)
            assert np.array_equal(got[i, :, :, :widths[i]], ref), (i, box, int(np.abs(got[i, :, :, :widths[i]].astype(int) - ref).max()))
            assert inv[i] == int(ref.max())
    with pytest.raises(ValueError):
        lineprep.prepare_box_lines(m, page, [(0, 0, pw + 1, 10)], pad=pad)
    with pytest.raises(Exception):
        lineprep.prepare_box_lines(m, page, [(0, 0, 1, 400)], pad=pad)    # int(1 * out_h / 400) == 0: Pillow raises too


@pytest.mark.gpu
def test_gpu_page_boxes_to_labels_equal_reference_transforms_plus_oracle():
    import vgsl_oracle as vo
    rng = np.random.default_rng(11)
    ph, pw = 600, 1500
    page = (rng.random((ph, pw)) * 255).astype(np.uint8)
    boxes = _random_boxes(rng, ph, pw, 20, hmin=30, hmax=110, wmin=200)
    om = vo.OracleModel(CFG2)
    wts = om.init_like_reference(4)
    m = kb.TorchVGSLModel(vgsl=CFG2)
    m.load_state_dict(wts)
    rec = kb.TorchSeqRecognizer(m, device='cuda:0')
    got = list(lineprep.recognize_boxes(rec, page, boxes, pad=16, batch_size=8))
    # reference: PIL steps -> PILToTensor -> ToDtype(scale) -> tensor_invert (dataset/utils.py:146-151), batches as rpred.py:126-131
    from kraken_b200.rpred import pad_batch
    k = 0
    for bi in range(0, len(boxes), 8):
        ts = []
        for box in boxes[bi:bi + 8]:
            u8 = torch.from_numpy(_pil_line(page, box, 48, 16))
            im = u8.to(torch.float32).mul_(1.0 / 255)
            ts.append(im.max() - im)
        seqs, lens = pad_batch(ts)
        _, _, _, ref_dec = vo.rec_predict(om, seqs, lens)
        r = got[k]; k += 1
        for i, d in enumerate(ref_dec):
            c = int(r['counts'][i])
            assert c == len(d)
            assert [(int(r['labels'][i, j]), int(r['starts'][i, j]), int(r['ends'][i, j])) for j in range(c)] == [(t[0], t[1], t[2]) for t in d]
