"""Stand-alone GPU diagnostic: the clustered tcgen05 recurrence of cfg2 with two / three groups of 8 lines per cluster (KB_LSTM_NG; unset =
the default of a synchronous call): device time of the stage, logits error against the CUDA-core recurrence, and the KB_LSTM_DBG
timeline of the default choice.
Usage: python tools/rec_ab.py 2> log"""
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
os.environ['KB_LSTM_TC'] = '0'
ref, _ = m.nn(x, lens)
os.environ.pop('KB_LSTM_TC')
for mode in ('', '2', '3', '', '2', '3'):
    if mode:
        os.environ['KB_LSTM_NG'] = mode
    else:
        os.environ.pop('KB_LSTM_NG', None)
    out, _ = m.nn(x, lens)
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
    print(f'groups per cluster {mode or "default"}: rec {acc["L_5.rec"]:.4f} ms, sum {sum(acc.values()):.4f} ms, logits vs CUDA-core recurrence {err:.2e}', file=sys.stderr)
os.environ.pop('KB_LSTM_NG', None)
os.environ['KB_LSTM_DBG'] = '1'
os.environ['KB_DEBUG'] = '1'
rec._recognize_raw(x, lens, want_probs=False)
