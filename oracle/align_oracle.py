"""
CPU oracle of the forced-alignment consumer of `return_logits` (second half).

TEST INFRASTRUCTURE ONLY - nothing under kraken_b200/ imports this (tests/test_abi.py enforces it).

Restates, per line, what `ForcedAlignmentTaskModel.predict` does with a record's `logits` (kraken/tasks/align.py:111-137):

    emission = record.logits.squeeze().log_softmax(0).T        align.py:119   (the "logits" are the softmax probabilities (C, T) of
                                                               kraken/lib/vgsl/rpred.py:226-227,200 - the second softmax is the reference's)
    trellis  = get_trellis(emission, labels)                   align.py:170-191
    path     = backtrack(trellis, emission, labels)            align.py:194-229
    segments = merge_repeats(path, text)                       align.py:232-249

as array code: float32 arithmetic where the reference's tensors are float32, Python floats (doubles) where the reference has left
torch (`.item()`, `sum(...) / n`).  Parity: PINNED - tests/test_align.py compares every function below with the reference's own
`get_trellis` / `backtrack` / `merge_repeats` (imported from /root/reference through oracle/refshim.py) on random and adversarial
emissions, bit for bit (trellis: torch.equal; path and segments: ==), and tests/golden/align_cases.npz holds outputs generated from the
reference by oracle/make_align_golden.py for the GPU box, where the reference does not exist.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

TOO_SHORT = -1          # align.py:113-117: fewer than 2 * len(labels) output frames -> an empty record
FAILED = -2             # align.py:228: ValueError("Failed to align")


def emission_from_probs(probs: torch.Tensor) -> torch.Tensor:
    """(C, T) probabilities -> (T, C) log-domain emission (align.py:119).  torch's own log_softmax: the oracle is bit-identical to
    the reference here by construction."""
    return probs.to(torch.float32).log_softmax(0).T


def trellis(emission: torch.Tensor, tokens: Sequence[int]) -> np.ndarray:
    """align.py:170-191.  (T + 1, J + 1) float32; row 0 / column 0 are the <SoS> paddings."""
    em = emission.numpy()
    T = em.shape[0]
    J = len(tokens)
    tok = np.asarray(tokens, dtype=np.int64)
    tr = np.empty((T + 1, J + 1), np.float32)
    tr[0, 0] = 0
    # torch.cumsum of float32 on the CPU accumulates in double and rounds every prefix (at::acc_type<float, false>)
    tr[1:, 0] = np.cumsum(em[:, 0].astype(np.float64)).astype(np.float32)
    tr[0, 1:] = -np.inf                                                   # trellis[0, -J:]  (J >= 1: align_line rejects J == 0)
    tr[T + 1 - J:, 0] = np.inf                                            # trellis[-J:, 0]
    for t in range(T):
        stay = tr[t, 1:] + em[t, 0]
        move = tr[t, :-1] + em[t, tok]
        tr[t + 1, 1:] = np.maximum(stay, move)
    return tr


def backtrack(tr: np.ndarray, emission: torch.Tensor, tokens: Sequence[int]):
    """align.py:194-229.  Returns [(token_index, time_index, score)] in time order, or None where the reference raises
    ValueError('Failed to align')."""
    em = emission.numpy()
    j = tr.shape[1] - 1
    t_start = int(np.argmax(tr[:, j]))                    # first maximum, as torch.argmax
    path = []
    for t in range(t_start, 0, -1):
        stayed = np.float32(tr[t - 1, j] + em[t - 1, 0])
        changed = np.float32(tr[t - 1, j - 1] + em[t - 1, tokens[j - 1]])
        moved = bool(changed > stayed)
        prob = float(torch.exp(emission[t - 1, tokens[j - 1] if moved else 0]))          # .exp().item(): torch's float32 exp
        path.append((j - 1, t - 1, prob))
        if moved:
            j -= 1
            if j == 0:
                return path[::-1]
    return None


def merge_repeats(path) -> List[Tuple[int, int, int, float]]:
    """align.py:232-249 with the token index in place of the character: [(token_index, start, end, score)], end exclusive."""
    out = []
    i1 = i2 = 0
    while i1 < len(path):
        while i2 < len(path) and path[i1][0] == path[i2][0]:
            i2 += 1
        score = sum(path[k][2] for k in range(i1, i2)) / (i2 - i1)
        out.append((path[i1][0], path[i1][1], path[i2 - 1][1] + 1, score))
        i1 = i2
    return out


def align_line(probs: torch.Tensor, tokens: Sequence[int]):
    """One record: (C, T) probabilities + label sequence -> (status, segments).  status = number of segments, TOO_SHORT or FAILED."""
    tokens = [int(t) for t in tokens]
    if len(tokens) == 0:
        raise IndexError('index -1 is out of bounds for dimension 0 with size 0')      # the reference's tokens[j - 1] on an empty tensor
    T = probs.shape[-1]
    if T < 2 * len(tokens):
        return TOO_SHORT, []
    em = emission_from_probs(probs)
    tr = trellis(em, tokens)
    path = backtrack(tr, em, tokens)
    if path is None:
        return FAILED, []
    segs = merge_repeats(path)
    return len(segs), segs


def scale_val(val: int, net_scale: float, in_scale: float, padding: int, max_val: int) -> int:
    """`_scale_val(val, 0, max_val)` (kraken/lib/vgsl/rpred.py:231) as align.py:131-132 calls it."""
    return int(round(min(max(((val * net_scale) - padding) * in_scale, 0), max_val - 1)))
