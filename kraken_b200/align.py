"""
Forced alignment of known transcriptions with the recogniser's output - the numeric part of the reference's
`ForcedAlignmentTaskModel.predict` (kraken/tasks/align.py:104-137) in one engine call per batch of lines: network -> probabilities ->
log-softmax emission -> trellis -> backtrack -> merged segments -> `_scale_val` positions, all on the device (`kb_forced_align`,
csrc/align.cuh).  The reference does this per line on the CPU from `record.logits`.

What stays with the caller, as in the reference: display-order conversion of the text (`get_display`, align.py:108), building the
`BaselineOCRRecord`, BiDi re-ordering (align.py:134-137).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from ._lib import check, lib
from .models import KrakenInputException, TorchSeqRecognizer
from .vgsl import _as_f32, _on_device, _ptr, _stream_for

__all__ = ['forced_align', 'forced_align_probs', 'align_records', 'TOO_SHORT', 'FAILED']

TOO_SHORT = -1
FAILED = -2


def forced_align(rec: TorchSeqRecognizer, line, lens, texts: Optional[Sequence[str]] = None,
                 labels: Optional[Sequence[Sequence[int]]] = None, orig_widths=None, padding: int = 16,
                 invert_max=None) -> List[Optional[list]]:
    """Aligns one batch.  `line` / `lens` / `invert_max`: as `TorchSeqRecognizer.recognize_records`.  `texts`: the transcriptions in
    display order, encoded with the recogniser's codec exactly as align.py:110 does (non-encodable code points are skipped unless the
    codec is strict); or `labels`: the label sequences themselves.

    Returns one entry per line: a list of (label, start, end, score) - `label` = `text[token_index]` as `merge_repeats` picks it
    (align.py:241; the token index itself when `labels` were given), start / end in output frames, or in pixels of the original line
    image when `orig_widths` is given (`_scale_val`, align.py:128-132), score = mean frame probability - or [] for a line with fewer
    than 2 * len(labels) output frames (the reference emits an empty record, align.py:113-117).  Raises ValueError('Failed to align')
    like the reference's backtrack, IndexError for an empty transcription."""
    net = rec.nn
    if (texts is None) == (labels is None):
        raise ValueError('give either texts or labels')
    if labels is None:
        if rec.codec is None:
            raise ValueError('aligning texts needs a codec')
        labels = [np.asarray(rec.codec.encode(t), dtype=np.int64).reshape(-1).tolist() for t in texts]
    is_u8 = isinstance(line, torch.Tensor) and line.dtype == torch.uint8 or isinstance(line, np.ndarray) and line.dtype == np.uint8
    x = (line if isinstance(line, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(line))).contiguous() if is_u8 else _as_f32(line)
    net._ensure_finalized(x)
    n, c, h, w = (int(v) for v in x.shape)
    if c != net.input[1]:
        raise ValueError(f'expected {net.input[1]} input channels, got {c}')
    if len(labels) != n:
        raise ValueError('one transcription per batch element')
    if _on_device(x) and x.device.index != net._device:
        x = x.to(f'cuda:{net._device}')
    for lab in labels:
        if len(lab) == 0:
            raise IndexError('index -1 is out of bounds for dimension 0 with size 0')          # tokens[j - 1] of an empty tensor, align.py:204
    tok_off = np.zeros(n + 1, np.int32)
    tok_off[1:] = np.cumsum([len(lab) for lab in labels])
    tokens = np.ascontiguousarray(np.concatenate([np.asarray(lab, np.int32).reshape(-1) for lab in labels]), dtype=np.int32)
    widths = np.ascontiguousarray(torch.as_tensor(lens).cpu().numpy(), dtype=np.int32) if lens is not None else np.full(n, w, np.int32)
    if widths.shape != (n,):
        raise ValueError('seq_lens must have one entry per batch element')
    ow = None
    if orig_widths is not None:
        ow = np.ascontiguousarray(np.asarray(orig_widths), dtype=np.int32)
        if ow.shape != (n,):
            raise ValueError('orig_widths must have one entry per batch element')
    inv = np.ascontiguousarray(np.asarray(invert_max), dtype=np.int16) if invert_max is not None else None
    dims = net.infer_dims(n, h, w)
    if dims[2] != 1:
        raise KrakenInputException('Expected dimension 3 to be 1, actual {}'.format(tuple(dims)))
    stride = int(max(len(lab) for lab in labels))
    seg_tok = np.zeros((n, stride), np.int32); seg_s = np.zeros((n, stride), np.int32); seg_e = np.zeros((n, stride), np.int32)
    seg_score = np.zeros((n, stride), np.float32); counts = np.zeros(n, np.int32); olens = np.zeros(n, np.int32)
    check(lib.kb_forced_align(net._h, _ptr(x), 1 if is_u8 else 0, int(_on_device(x)), n, h, w, widths.ctypes.data,
                              inv.ctypes.data if inv is not None else None, float(rec.temperature), tokens.ctypes.data, tok_off.ctypes.data,
                              ow.ctypes.data if ow is not None else None, int(padding), seg_tok.ctypes.data, seg_s.ctypes.data,
                              seg_e.ctypes.data, seg_score.ctypes.data, counts.ctypes.data, stride, olens.ctypes.data,
                              _stream_for(x, net._device)))
    out: List[Optional[list]] = []
    for i in range(n):
        k = int(counts[i])
        if k == FAILED:
            raise ValueError('Failed to align')
        if k == TOO_SHORT:
            out.append([])
            continue
        src = texts[i] if texts is not None else None
        out.append([(src[int(seg_tok[i, j])] if src is not None else int(seg_tok[i, j]), int(seg_s[i, j]), int(seg_e[i, j]), float(seg_score[i, j]))
                    for j in range(k)])
    return out


def forced_align_probs(probs, labels: Sequence[Sequence[int]], lens=None, device: int = 0) -> List[Optional[list]]:
    """The same from probabilities the caller already holds: `probs` (N, C, T) - the `logits` of records produced with
    `return_logits` (kraken/lib/vgsl/rpred.py:200), host or CUDA tensor - and one label sequence per line.  `lens`: valid frames per
    line.  Returns per line [(token_index, start, end, score)] in frames, [] for lines that are too short; raises like `forced_align`."""
    x = _as_f32(probs)
    if x.ndim == 2:
        x = x[None]
    n, c, t = (int(v) for v in x.shape)
    if len(labels) != n:
        raise ValueError('one label sequence per batch element')
    for lab in labels:
        if len(lab) == 0:
            raise IndexError('index -1 is out of bounds for dimension 0 with size 0')
    tok_off = np.zeros(n + 1, np.int32)
    tok_off[1:] = np.cumsum([len(lab) for lab in labels])
    tokens = np.ascontiguousarray(np.concatenate([np.asarray(lab, np.int32).reshape(-1) for lab in labels]), dtype=np.int32)
    ln = np.ascontiguousarray(torch.as_tensor(lens).cpu().numpy(), dtype=np.int32) if lens is not None else None
    if ln is not None and ln.shape != (n,):
        raise ValueError('lens must have one entry per batch element')
    stride = int(max(len(lab) for lab in labels))
    seg_tok = np.zeros((n, stride), np.int32); seg_s = np.zeros((n, stride), np.int32); seg_e = np.zeros((n, stride), np.int32)
    seg_score = np.zeros((n, stride), np.float32); counts = np.zeros(n, np.int32)
    dev = x.device.index if _on_device(x) else int(device)
    check(lib.kb_forced_align_probs(_ptr(x), int(_on_device(x)), n, c, t, ln.ctypes.data if ln is not None else None, tokens.ctypes.data,
                                    tok_off.ctypes.data, seg_tok.ctypes.data, seg_s.ctypes.data, seg_e.ctypes.data, seg_score.ctypes.data,
                                    counts.ctypes.data, stride, dev, _stream_for(x, dev)))
    out: List[Optional[list]] = []
    for i in range(n):
        k = int(counts[i])
        if k == FAILED:
            raise ValueError('Failed to align')
        out.append([] if k == TOO_SHORT else [(int(seg_tok[i, j]), int(seg_s[i, j]), int(seg_e[i, j]), float(seg_score[i, j])) for j in range(k)])
    return out


def align_records(logits: Sequence, labels: Sequence[Sequence[int]], device: int = 0) -> List[Optional[list]]:
    """Aligns the records of one page in ONE engine call: `logits[i]` is record i's `logits` - (C, T_i) or (C, 1, T_i), every record
    its own length, as kraken/lib/vgsl/rpred.py:200 cuts them - `labels[i]` its encoded transcription.  This is what replaces
    align.py:119-122 inside the reference's per-record loop."""
    ts = [torch.as_tensor(l).float().reshape(int(l.shape[0]), -1) for l in logits]
    if not ts:
        return []
    c = int(ts[0].shape[0])
    tmax = max(int(t.shape[1]) for t in ts)
    dev = next((t.device for t in ts if t.is_cuda), torch.device('cpu'))
    batch = torch.zeros((len(ts), c, tmax), dtype=torch.float32, device=dev)
    for i, t in enumerate(ts):
        if int(t.shape[0]) != c:
            raise ValueError('records of different models in one call')
        batch[i, :, :t.shape[1]] = t.to(dev)
    return forced_align_probs(batch, labels, lens=[int(t.shape[1]) for t in ts], device=device)
