"""
Segmentation-network forward of blla (reference: kraken/lib/vgsl/spred.py:237-287 `_compute_segmentation_map`,
kraken/blla.py:59-141 `compute_segmentation_map`): net -> nearest upsample to the scaled page size -> sigmoid ->
crop padding.  Everything after the heat map (vectorisation, reading order) is CPU geometry and out of scope.
Unlike the reference (batch 1, spred.py:268) pages of equal size can be batched.
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import numpy as np
import torch

from ._lib import check, lib
from .vgsl import TorchVGSLModel, _as_f32, _on_device, _ptr, _stream_for

__all__ = ['compute_segmentation_map', 'segmentation_heatmap']


def segmentation_heatmap(model: TorchVGSLModel, pages, size: Optional[Sequence[int]] = None):
    """pages: (N, C, H, W) tensor already transformed to network input.  Returns sigmoid heat maps
    (N, C', size[0], size[1]) - on the GPU if `pages` is a CUDA tensor, else a CPU tensor (a pinned buffer owned by the model and
    reused by the next call with the same shape: copy it if you keep it)."""
    model._ensure_finalized(pages)
    x = _as_f32(pages)
    if x.ndim == 3:
        x = x[None]
    n, c, h, w = (int(v) for v in x.shape)
    if c != model.input[1]:
        raise ValueError(f'expected {model.input[1]} input channels, got {c}')
    if _on_device(x) and x.device.index != model._device:
        x = x.to(f'cuda:{model._device}')
    if size is None:
        size = (h, w)
    oc = model.infer_dims(n, h, w)[1]
    on_dev = _on_device(x)
    shape = (n, oc, int(size[0]), int(size[1]))
    if on_dev:
        out = torch.empty(shape, dtype=torch.float32, device=x.device)
    else:
        # host result: a PINNED buffer (cached per shape on the model) - a pageable destination made the 544 MB heat-map read-back of
        # a cfg3 batch crawl at ~4.7 GB/s, 6.5x longer than the forward itself (round-1 measurement)
        cache = model.__dict__.setdefault('_pinned_out', {})
        out = cache.get(shape)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32)
            if torch.cuda.is_available():
                out = out.pin_memory()
            cache.clear()
            cache[shape] = out
    check(lib.kb_segment(model._h, _ptr(x), int(on_dev), n, h, w, int(size[0]), int(size[1]), out.data_ptr(), int(on_dev), _stream_for(x, model._device)))
    return out


def compute_segmentation_map(model: TorchVGSLModel, tensor_im, scal_shape: Optional[Sequence[int]] = None,
                             padding: Union[int, Sequence[int]] = 0) -> dict:
    """Mirror of the dict the reference returns (minus the PIL-side `scal_im`): heat map as numpy (C', H, W),
    class map and bounding regions from the model metadata, padding removed (spred.py:271-287)."""
    if isinstance(padding, int):
        padding = (padding,) * 4
    elif len(padding) == 2:
        padding = (padding[0], padding[0], padding[1], padding[1])
    t = tensor_im if tensor_im.ndim == 4 else tensor_im[None]
    size = tuple(scal_shape) if scal_shape is not None else tuple(t.shape[2:])
    o = segmentation_heatmap(model, t, size)
    pad = [p if p else None for p in padding]
    pad[1] = -pad[1] if pad[1] else None
    pad[3] = -pad[3] if pad[3] else None
    o = o[:, :, pad[2]:pad[3], pad[0]:pad[1]]
    hm = o.squeeze().cpu().float().numpy()
    if not o.is_cuda:
        hm = hm.copy()                 # the host result is a view of the model's reusable pinned buffer
    return {'heatmap': hm,
            'cls_map': model.user_metadata.get('class_mapping'),
            'bounding_regions': model.user_metadata.get('bounding_regions', None)}
