"""
Batched recognition over already-extracted line tensors (reference: kraken/lib/vgsl/rpred.py:126-131,171-176
batching; kraken/rpred.py:61-182,373-391 tag -> model routing of `mm_rpred`).

Line extraction / PIL transforms (extract_polygons, ImageInputTransforms) stay on the CPU side of kraken and are
out of scope; these helpers start where the reference has `(C, H, W_i)` tensors in its input queue:
  * right-pad with 0 to the batch maximum, stack, seq_lens = widths  - identical padded-batch semantics, arrival order
  * one fused engine call per batch
  * per-line routing to a model by tag (mm_rpred), batches formed per model
  * `_scale_val`-style position scaling is left to the caller's record construction
For the legacy iterator API itself, hand a `kraken_b200.models.TorchSeqRecognizer` to the reference's own
`kraken.rpred.rpred/mm_rpred` - it is duck-type compatible.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Optional, Sequence

import numpy as np
import torch

from .models import TorchSeqRecognizer
from .vgsl import TorchVGSLModel

__all__ = ['pad_batch', 'recognize_lines', 'mm_recognize_lines', 'resolve_type_to_model']


def pad_batch(lines: Sequence[torch.Tensor]):
    """[(C,H,W_i)] -> ((N,C,H,Wmax) zero right-padded, LongTensor widths) exactly like rpred.py:129-131."""
    max_len = max(int(l.shape[2]) for l in lines)
    seqs = torch.stack([torch.nn.functional.pad(l, pad=(0, max_len - int(l.shape[2]))) for l in lines])
    return seqs, torch.LongTensor([int(l.shape[2]) for l in lines])


def recognize_lines(model, lines: Sequence[torch.Tensor], batch_size: int = 64, temperature: float = 1.0):
    """Runs `lines` through `model` (TorchVGSLModel or TorchSeqRecognizer) in arrival order, `batch_size` at a time.
    Lines that are empty/constant yield empty results (rpred.py:104-113).  Returns a list (one per line) of
    [(char, start, end, conf)] when the model has a codec, else label tuples."""
    rec = model if isinstance(model, TorchSeqRecognizer) else TorchSeqRecognizer(model, temperature=temperature, device=None)
    results: list = [None] * len(lines)
    queue = []
    for idx, l in enumerate(lines):
        if l is None or 0 in l.shape or float(l.max()) == float(l.min()):
            results[idx] = []
        else:
            queue.append((idx, l))
    for i in range(0, len(queue), batch_size):
        chunk = queue[i:i + batch_size]
        seqs, lens = pad_batch([l for _, l in chunk])
        dec = rec.predict(seqs, lens) if rec.codec is not None else rec.predict_labels(seqs, lens)
        for (idx, _), d in zip(chunk, dec):
            results[idx] = d
    return results


def resolve_type_to_model(tag: Optional[str], model_map: dict, default=None):
    """kraken/rpred.py:373-391."""
    if not tag and default is not None:
        return 'default', default
    if tag in model_map:
        return tag, model_map[tag]
    if tag and default is not None:
        return tag, default
    raise KeyError(f'No model for type {tag}')


def mm_recognize_lines(nets: dict, lines: Sequence[torch.Tensor], tags: Sequence[Optional[str]], batch_size: int = 64,
                       tags_ignore: Optional[Sequence[str]] = None):
    """Multi-model recognition: every line goes to the model its tag selects (a defaultdict supplies the default);
    ignored tags produce empty results.  Batches are formed per model, results returned in input order."""
    default = nets.default_factory() if isinstance(nets, defaultdict) and nets.default_factory else None
    groups: dict = {}
    results: list = [None] * len(lines)
    for idx, (l, tag) in enumerate(zip(lines, tags)):
        if tags_ignore and tag in tags_ignore:
            results[idx] = []
            continue
        key, net = resolve_type_to_model(tag, nets, default)
        groups.setdefault(id(net), (net, []))[1].append(idx)
    for net, idxs in groups.values():
        out = recognize_lines(net, [lines[i] for i in idxs], batch_size)
        for i, o in zip(idxs, out):
            results[i] = o
    return results
