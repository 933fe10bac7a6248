"""
GPU line pre-processing for bbox lines: page image + bounding boxes -> the uint8 line batch the
recogniser's `recognize_u8` / `submit` take, on the device.

Replaces (bit for bit, see csrc/line_prep.cuh and tests/test_line_prep.py):
  `im.crop(box)`                                    kraken/lib/segmentation.py:1631-1643   (bbox lines, horizontal text)
  Grayscale -> pil_fixed_resize (LANCZOS) -> Pad    kraken/lib/dataset/utils.py:123-147, kraken/lib/functional_im_transforms.py:58-82
and hands over to the device half that already existed (`kb_recognize_u8`: ToDtype(scale), tensor_invert, batch zero padding).

Stays in kraken: baseline / polygon extraction (`extract_polygons` with baselines), the centre normaliser of legacy bbox models
(`valid_norm=True`), vertical text (the 90 degree rotation), forced binarisation.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from ._lib import check, lib
from .models import TorchSeqRecognizer
from .vgsl import TorchVGSLModel

__all__ = ['line_width', 'prepare_box_lines', 'recognize_boxes']


def line_width(box_w: int, box_h: int, out_h: int, pad: int) -> int:
    """Padded width of a prepared line: int(w * out_h / h) + 2 * pad; 0 where the reference's resize would raise."""
    return int(lib.kb_line_width(int(box_w), int(box_h), int(out_h), int(pad)))


def _page_array(page):
    """PIL image / ndarray / tensor -> (uint8 array or cuda tensor, H, W, channels) in the layouts PIL uses ('L': HxW, 'RGB': HxWx3)."""
    if isinstance(page, torch.Tensor):
        if page.dtype != torch.uint8 or page.ndim not in (2, 3):
            raise ValueError('page tensor must be uint8, (H, W) or (H, W, 3)')
        page = page.contiguous()
        ch = 1 if page.ndim == 2 else int(page.shape[2])
        return page, int(page.shape[0]), int(page.shape[1]), ch
    if hasattr(page, 'mode') and hasattr(page, 'convert'):          # PIL image: same conversions as the reference's mode_transform
        if page.mode not in ('L', 'RGB'):
            page = page.convert('L' if page.mode in ('1', 'I', 'F', 'LA', 'P') else 'RGB')
        page = np.asarray(page)
    a = np.ascontiguousarray(page)
    if a.dtype != np.uint8 or a.ndim not in (2, 3):
        raise ValueError('page must be uint8, (H, W) or (H, W, 3)')
    return a, int(a.shape[0]), int(a.shape[1]), 1 if a.ndim == 2 else int(a.shape[2])


def prepare_box_lines(model: TorchVGSLModel, page, boxes: Sequence[Sequence[int]], pad: int = 16, out_h: Optional[int] = None):
    """Crops `boxes` = [(x0, y0, x1, y1)] out of `page`, converts to grayscale, scales every crop to the model's input height with
    Pillow's LANCZOS arithmetic and pads `pad` white columns on both ends - on the model's GPU.  Returns
    (lines: cuda uint8 tensor (n, 1, out_h, wmax), widths: int32 array, invert_max: int16 array) ready for
    `TorchSeqRecognizer.recognize_u8(lines, widths, invert_max)` / `.submit(lines, widths, invert_max)`."""
    net = model.nn if isinstance(model, TorchSeqRecognizer) else model
    net._ensure_finalized(None)
    if out_h is None:
        out_h = int(net.input[2])
    if out_h <= 0:
        raise ValueError('the model has a variable input height: pass out_h')
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.int32).reshape(-1, 4))
    n = int(b.shape[0])
    if n == 0:
        raise ValueError('no boxes')
    arr, ph, pw, ch = _page_array(page)
    ws = [line_width(int(x1 - x0), int(y1 - y0), out_h, pad) for x0, y0, x1, y1 in b.tolist()]
    if min(ws) - 2 * pad < 1:
        raise ValueError('height and width must be > 0')
    wmax = max(ws)
    dev = torch.device(f'cuda:{net._device}')
    lines = torch.empty((n, 1, out_h, wmax), dtype=torch.uint8, device=dev)
    widths = np.zeros(n, np.int32)
    inv = np.zeros(n, np.int16)
    on_dev = isinstance(arr, torch.Tensor) and arr.is_cuda
    if on_dev and arr.device.index != net._device:                         # a page resident on another GPU: bring it to the model's
        arr = arr.to(dev)
    if isinstance(arr, torch.Tensor) and not arr.is_cuda:
        arr = arr.numpy()
    ptr = arr.data_ptr() if on_dev else arr.ctypes.data
    stream = torch.cuda.current_stream(dev).cuda_stream
    check(lib.kb_prepare_lines_u8(net._h, ptr, int(on_dev), ph, pw, ch, n, b.ctypes.data, int(out_h), int(pad), lines.data_ptr(), wmax,
                                  widths.ctypes.data, inv.ctypes.data, stream))
    return lines, widths, inv


def recognize_boxes(rec: TorchSeqRecognizer, page, boxes: Sequence[Sequence[int]], pad: int = 16, batch_size: int = 64):
    """bbox lines of one page straight to decoded label blocks: crop / grayscale / LANCZOS / pad / scale / invert / net / CTC decode all on
    the GPU.  Yields one `recognize_u8` result dict per batch of `batch_size` boxes (arrival order, like rpred.py:126-131)."""
    b = np.asarray(boxes, dtype=np.int32).reshape(-1, 4)
    arr, ph, pw, ch = _page_array(page)
    if not (isinstance(arr, torch.Tensor) and arr.is_cuda):                 # one upload of the page for all its batches
        arr = torch.as_tensor(arr).to(f'cuda:{rec.nn._device}')
    for i in range(0, b.shape[0], batch_size):
        lines, widths, inv = prepare_box_lines(rec, arr, b[i:i + batch_size], pad)
        yield rec.recognize_u8(lines, widths, inv)
