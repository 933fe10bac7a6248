"""Generates tests/golden/align_cases.npz FROM THE UNMODIFIED REFERENCE (kraken/tasks/align.py get_trellis / backtrack / merge_repeats,
imported through oracle/refshim.py): probabilities, label sequences and the reference's segments for the forced-alignment tests on
the GPU box, where /root/reference does not exist.  TEST INFRASTRUCTURE ONLY.  Usage: python oracle/make_align_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import refshim

refshim.install()
from kraken.tasks import align as ra          # noqa: E402

from test_align import random_case            # noqa: E402  (the same generator the tests use)


def main():
    rng = np.random.default_rng(1234)
    out = {}
    shapes = [(40, 64, 9), (200, 200, 60), (200, 300, 150), (30, 50, 25), (12, 9, 5), (97, 157, 1), (200, 500, 90), (5, 40, 20), (64, 120, 33)]
    k = 0
    for C, T, J in shapes:
        for peaky in (True, False):
            p, tokens = random_case(rng, C, T, J, peaky=peaky)
            out[f'probs_{k}'] = p.numpy()
            out[f'tokens_{k}'] = np.asarray(tokens, np.int32)
            labels = torch.tensor(tokens).long()
            if p.shape[-1] < 2 * len(labels):                                   # align.py:113-117
                out[f'status_{k}'] = np.int32(-1)
            else:
                em = p.squeeze().log_softmax(0).T
                tr = ra.get_trellis(em, labels)
                try:
                    path = ra.backtrack(tr, em, labels)
                    segs = ra.merge_repeats(path, list(range(len(tokens))))
                    out[f'status_{k}'] = np.int32(len(segs))
                    out[f'seg_token_{k}'] = np.asarray([s.label for s in segs], np.int32)
                    out[f'seg_start_{k}'] = np.asarray([s.start for s in segs], np.int32)
                    out[f'seg_end_{k}'] = np.asarray([s.end for s in segs], np.int32)
                    out[f'seg_score_{k}'] = np.asarray([s.score for s in segs], np.float64)
                except ValueError:
                    out[f'status_{k}'] = np.int32(-2)
            k += 1
    out['n_cases'] = np.int32(k)
    path = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'align_cases.npz')
    np.savez_compressed(path, **out)
    print(path, k, 'cases', os.path.getsize(path), 'bytes; statuses', [int(out[f'status_{i}']) for i in range(k)])


if __name__ == '__main__':
    main()
