"""Stand-alone GPU diagnostic: device time of the forced-alignment stage (csrc/align.cuh) on a cfg2 batch, next to the oracle's CPU time
for the same lines.  Usage: python tools/align_time.py 2> log"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import torch

import kraken_b200 as kb
from kraken_b200 import align

CFG2 = '[1,48,0,1 Cr3,3,32 Mp2,2 Cr3,3,64 Mp2,2 S1(1x0)1,3 Lbx256 O1c200]'
torch.manual_seed(0)
m = kb.TorchVGSLModel(vgsl=CFG2)
m.init_weights()
rec = kb.TorchSeqRecognizer(m, device='cuda:0')
for n, w, j in ((64, 800, 40), (64, 800, 95), (64, 2000, 120)):
    x = torch.rand(n, 1, 48, w).cuda()
    lens = torch.full((n,), w)
    rng = np.random.default_rng(0)
    labels = [rng.integers(1, 200, j).tolist() for _ in range(n)]
    for _ in range(3):
        out = align.forced_align(rec, x, lens, labels=labels)
    m.set_timing(True)
    acc = {}
    t0 = time.perf_counter()
    for _ in range(10):
        align.forced_align(rec, x, lens, labels=labels)
        for k, v in m.last_timing():
            acc[k] = acc.get(k, 0.0) + v / 10
    wall = (time.perf_counter() - t0) / 10
    m.set_timing(False)
    print(f'{n} lines 48x{w}, {j} labels each: align stage {acc.get("align", float("nan")):.3f} ms, whole call {wall * 1e3:.2f} ms wall, '
          f'network stages {sum(v for k, v in acc.items() if k != "align"):.3f} ms', file=sys.stderr)
    # the reference's way: probabilities to the host, per line log_softmax + Python trellis loop (oracle = the reference's arithmetic)
    import align_oracle as ao
    probs = rec.predict_probs(x, lens) if hasattr(rec, 'predict_probs') else None
    if probs is None:
        rec.keep_outputs = True
        rec.predict_labels(x, lens)
        probs = torch.as_tensor(rec.outputs).float().cpu()
        rec.keep_outputs = False
    t0 = time.perf_counter()
    for i in range(8):
        ao.align_line(probs[i], labels[i])
    print(f'    oracle (numpy restatement of align.py) on the host: {(time.perf_counter() - t0) / 8 * 1e3:.1f} ms per line', file=sys.stderr)
