"""Stand-alone GPU diagnostic: the clustered recurrence of cfg2 with 2 / 3 groups per cluster (KB_LSTM_NG), with and without the
hand-over of the tensor pipe between groups (KB_LSTM_ALT) and under the KB_LSTM_DBG bits (1: clock64 timeline of steps 100..103, 2:
issuer polls its h barriers with test_wait instead of try_wait, 4: epilogue polls mma_done); device time of the stage and the logits
against the first run per setting.  Usage: python tools/rec_dbg.py 2> log"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import kraken_b200 as kb

CFG2 = '[1,48,0,1 Cr3,3,32 Mp2,2 Cr3,3,64 Mp2,2 S1(1x0)1,3 Lbx256 O1c200]'
torch.manual_seed(0)
m = kb.TorchVGSLModel(vgsl=CFG2)
m.init_weights()
rec = kb.TorchSeqRecognizer(m, device='cuda:0')
x = torch.rand(64, 1, 48, 800).cuda()
lens = torch.full((64,), 800)
import numpy as np
ref = None
for ng, alt, mode in (('2', '1', '0'), ('3', '0', '0'), ('4', '0', '0'), ('4', '1', '0'), ('2', '1', '0'), ('3', '0', '0'), ('4', '0', '0')):
    os.environ['KB_LSTM_NG'] = ng
    os.environ['KB_LSTM_ALT'] = alt
    os.environ['KB_LSTM_DBG'] = mode
    out, _ = m.nn(x, lens)
    if ref is None:
        ref = out
    err = float((out - ref).abs().max() / ref.abs().max())
    for _ in range(3):
        rec._recognize_raw(x, lens, want_probs=False)
    m.set_timing(True)
    acc = {}
    for _ in range(20):
        rec._recognize_raw(x, lens, want_probs=False)
        for k, v in m.last_timing():
            acc[k] = acc.get(k, 0.0) + v / 20
    m.set_timing(False)
    print(f'NG={ng} ALT={alt} KB_LSTM_DBG={mode}: rec {acc["L_5.rec"]:.4f} ms, logits vs the first run {err:.2e}', file=sys.stderr)
os.environ['KB_LSTM_DBG'] = '1'
rec._recognize_raw(x, lens, want_probs=False)
