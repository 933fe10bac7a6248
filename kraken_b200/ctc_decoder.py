"""
GPU replacement of kraken.lib.ctc_decoder.greedy_decoder (reference: kraken/lib/ctc_decoder.py:35-72),
usable as the `decoder` hook of RecognitionInferenceConfig / TorchSeqRecognizer.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import check, lib

__all__ = ['greedy_decoder', 'unpack_decoded']


def unpack_decoded(labels, starts, ends, confs, counts):
    """fixed-stride engine output -> list[list[(label, start, end, conf)]] of python scalars."""
    out = []
    for i, c in enumerate(counts.tolist()):
        out.append(list(zip(labels[i, :c].tolist(), starts[i, :c].tolist(), ends[i, :c].tolist(), confs[i, :c].tolist())))
    return out


def greedy_decoder(outputs, seq_lens=None, device: int | None = None):
    """
    Best-path CTC decoding of a (C, W) or (N, C, W) softmax tensor.

    Returns a list (one per line) of tuples (class, start, end, max) exactly like the reference: runs of
    equal arg-max labels, blank (0) dropped, `max` the highest probability inside the run.
    """
    t = outputs if isinstance(outputs, torch.Tensor) else torch.as_tensor(np.asarray(outputs))
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3:
        raise ValueError('outputs must be (C, W) or (N, C, W)')
    n, c, w = (int(v) for v in t.shape)
    if n == 1 and seq_lens is None:
        lens = np.array([w], dtype=np.int32)
    elif seq_lens is None:
        raise ValueError('seq_lens need to be set for batch decoding.')
    else:
        lens = np.ascontiguousarray(torch.as_tensor(seq_lens).cpu().numpy(), dtype=np.int32)
    t = t.detach().to(torch.float32).contiguous()
    on_dev = t.is_cuda
    dev = t.device.index if on_dev else (device if device is not None else 0)
    stride = max(w, 1)
    labels = np.zeros((n, stride), np.int32)
    starts = np.zeros((n, stride), np.int32)
    ends = np.zeros((n, stride), np.int32)
    confs = np.zeros((n, stride), np.float32)
    counts = np.zeros(n, np.int32)
    stream = torch.cuda.current_stream(t.device).cuda_stream if on_dev else None
    check(lib.kb_ctc_greedy_decode(t.data_ptr(), int(on_dev), n, c, w, lens.ctypes.data, labels.ctypes.data, starts.ctypes.data,
                                   ends.ctypes.data, confs.ctypes.data, counts.ctypes.data, stride, dev, stream))
    return unpack_decoded(labels, starts, ends, confs, counts)
