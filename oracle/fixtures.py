"""
oracle/fixtures.py - helpers shared by the fixture generator (oracle/make_golden.py) and the tests that read the
committed fixtures of the TRAINED models (tests/golden/trained_*.npz).  TEST INFRASTRUCTURE ONLY.
"""
import numpy as np
import torch


def line_from_u8(u8: np.ndarray) -> torch.Tensor:
    """(1, H, W) uint8 pixels -> the (1, 1, H, W) float tensor the reference's ImageInputTransforms hands to the net:
    v2.ToDtype(float32, scale=True) then tensor_invert (kraken/lib/dataset/utils.py:148-151,
    kraken/lib/functional_im_transforms.py:58-59).  make_golden.py proves this reproduces the captured input bit for bit."""
    im = torch.from_numpy(np.ascontiguousarray(u8)).to(torch.float32).mul_(1.0 / 255)
    return (im.max() - im)[None]


def trained_lines(g) -> list:
    return [line_from_u8(g[f'u8::{i}']) for i in range(int(g['n_lines']))]


def trained_weights(g) -> dict:
    return {k[3:]: torch.from_numpy(np.asarray(g[k], np.float32)) for k in g if k.startswith('w::')}


def trained_expected(g, i):
    """[(label, start, end, conf)] of line i as the reference decoded it."""
    lab, st, en, cf = g[f'dec_label::{i}'], g[f'dec_start::{i}'], g[f'dec_end::{i}'], g[f'dec_conf::{i}']
    n = int(g[f'dec_count::{i}'])
    return [(int(lab[j]), int(st[j]), int(en[j]), float(cf[j])) for j in range(n)]


def blla_page_tensor(path, size=(1350, 1800)) -> torch.Tensor:
    """A 2400x3200-class page as the segmentation net sees it (BASELINE cfg3): RGB, resized to 1350 x 1800 (W x H) with LANCZOS,
    scaled to [0, 1] (kraken/lib/vgsl/spred.py:252-266 without the padding)."""
    from PIL import Image
    im = Image.open(path).convert('RGB').resize(size, Image.LANCZOS)
    a = np.asarray(im, np.uint8).transpose(2, 0, 1).copy()
    return torch.from_numpy(a).to(torch.float32).mul_(1.0 / 255)[None]
