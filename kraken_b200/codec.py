"""
Label <-> code point codec used right after CTC decoding (reference: kraken/lib/codec.py:36-195).
Host-side only; kept semantically identical for `decode` (the only method on the rpred path) and
`encode`, with a vectorised fast path for the usual 1:1 label tables.
"""
from __future__ import annotations

from collections import Counter
from typing import Sequence, Union

import numpy as np

__all__ = ['PytorchCodec', 'KrakenCodecException', 'KrakenEncodeException']


class KrakenCodecException(Exception):
    pass


class KrakenEncodeException(Exception):
    pass


class PytorchCodec:
    def __init__(self, charset: Union[dict, Sequence[str], str], strict: bool = False):
        if isinstance(charset, dict):
            self.c2l = {k: list(v) for k, v in charset.items()}
        else:
            counts = Counter(charset)
            if len(counts) < len(charset):
                raise KrakenCodecException(f'Duplicate entry in codec definition string: {counts}')
            self.c2l = {k: [v] for v, k in enumerate(sorted(charset), start=1)}
        self.c_sorted = sorted(self.c2l.keys(), key=len, reverse=True)
        self.l2c = {tuple(v): k for k, v in self.c2l.items()}
        self.l2c_single = {k[0]: v for k, v in self.l2c.items() if len(k) == 1}
        self.strict = strict
        if not self.is_valid:
            raise KrakenCodecException('Codec is not valid (non-singular/non-prefix free).')

    def __len__(self) -> int:
        return len(self.l2c)

    @property
    def is_valid(self) -> bool:
        if len(self.l2c) != len(self.c2l):
            return False
        codes = sorted(self.l2c.keys())
        # prefix-freeness: after sorting, a prefix sorts directly before some extension of it
        for a, b in zip(codes, codes[1:]):
            if b[:len(a)] == a:
                return False
        return True

    @property
    def max_label(self) -> int:
        return max(l for ls in self.c2l.values() for l in ls)

    def encode(self, s: str):
        import torch
        labels: list[int] = []
        i = 0
        multi = [c for c in self.c_sorted if len(c) > 1]
        while i < len(s):
            for code in multi:
                if s.startswith(code, i):
                    labels.extend(self.c2l[code])
                    i += len(code)
                    break
            else:
                if s[i] in self.c2l:
                    labels.extend(self.c2l[s[i]])
                elif self.strict:
                    raise KrakenEncodeException(f'Non-encodable sequence {s[i:i + 5]}... encountered.')
                i += 1
        return torch.IntTensor(labels)

    def decode(self, labels: Sequence[tuple[int, int, int, float]]) -> list[tuple[str, int, int, float]]:
        """(label, start, end, conf)* -> (code point, start, end, conf)*; multi-label codes aggregate
        min start / max end / mean confidence (codec.py:148-195)."""
        labs = tuple(int(l) for l, *_ in labels)
        out = []
        i = 0
        n = len(labs)
        while i < n:
            lab = labs[i]
            if lab in self.l2c_single:
                _, s, e, c = labels[i]
                out.extend((ch, s, e, c) for ch in self.l2c_single[lab])
                i += 1
                continue
            for code, chars in self.l2c.items():
                k = len(code)
                if code == labs[i:i + k]:
                    s = labels[i][1]
                    e = labels[i + k - 1][2]
                    c = np.mean([labels[j][3] for j in range(i, i + k)])
                    out.extend((ch, s, e, c) for ch in chars)
                    i += k
                    break
            else:
                if self.strict:
                    raise KrakenEncodeException(f'Non-decodable sequence {labs[i:i + 5]}... encountered.')
                i += 1
        return out
