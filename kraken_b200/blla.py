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

__all__ = ['compute_segmentation_map', 'segmentation_heatmap', 'apply_legacy_mask']


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


def apply_legacy_mask(tensor_im: torch.Tensor, mask_tensor: torch.Tensor) -> torch.Tensor:
    """The legacy `mask` argument of `kraken.blla.compute_segmentation_map` (kraken/blla.py:107-116): the mask image goes through the
    same ImageInputTransforms as the page (hence arrives INVERTED, like the page) and `tensor_im[~transforms(mask).bool()] = 0` zeroes
    the page wherever the transformed mask is 0.  `tensor_im`: (C, H, W) network input, `mask_tensor`: the transformed mask with the
    shape the reference's boolean indexing accepts (equal to `tensor_im`'s, or (1, H, W) for one-channel models).  Stays on the device
    the page tensor is on; returns a new tensor."""
    if tuple(mask_tensor.shape) != tuple(tensor_im.shape):
        # the reference's boolean index raises for a (1, H, W) mask on a 3-channel page
        raise IndexError(f'The shape of the mask {list(mask_tensor.shape)} at index 0 does not match the shape of the indexed tensor '
                         f'{list(tensor_im.shape)} at index 0')
    keep = mask_tensor.to(tensor_im.device).bool()
    return torch.where(keep, tensor_im, torch.zeros((), dtype=tensor_im.dtype, device=tensor_im.device))


def compute_segmentation_map(model: TorchVGSLModel, tensor_im, scal_shape: Optional[Sequence[int]] = None,
                             padding: Union[int, Sequence[int]] = 0, mask_tensor: Optional[torch.Tensor] = None) -> dict:
    """Mirror of the dict the reference returns (minus the PIL-side `scal_im`): heat map as numpy (C', H, W),
    class map and bounding regions from the model metadata, padding removed (spred.py:271-287).  `mask_tensor`: the legacy masking of
    `kraken.blla.compute_segmentation_map` (blla.py:107-116), see `apply_legacy_mask` (single page only, as in the reference)."""
    if isinstance(padding, int):
        padding = (padding,) * 4
    elif len(padding) == 2:
        padding = (padding[0], padding[0], padding[1], padding[1])
    if mask_tensor is not None:
        if tensor_im.ndim == 4:
            if tensor_im.shape[0] != 1:
                raise ValueError('the legacy mask applies to a single page')
            tensor_im = apply_legacy_mask(tensor_im[0], mask_tensor)[None]
        else:
            tensor_im = apply_legacy_mask(tensor_im, mask_tensor)
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
