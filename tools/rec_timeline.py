"""Stand-alone GPU diagnostic (not a pytest module): one cfg2 batch (64 x 48 x 800) with KB_LSTM_DBG=1 - the clustered tcgen05
recurrence prints the clock64 timeline of steps 100..103 of cluster 0 to stderr - plus the per-stage device times of a few calls.
Usage under gpurun:  python tools/rec_timeline.py 2> gpurun_out/rec_timeline.log"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch

import kraken_b200 as kb
import vgsl_oracle as vo

CFG2 = '[1,48,0,1 Cr3,3,32 Mp2,2 Cr3,3,64 Mp2,2 S1(1x0)1,3 Lbx256 O1c200]'
om = vo.OracleModel(CFG2)
m = kb.TorchVGSLModel(vgsl=CFG2)
m.load_state_dict(om.init_like_reference(0))
rec = kb.TorchSeqRecognizer(m, device='cuda:0')
x = torch.rand(64, 1, 48, 800).cuda()
lens = torch.full((64,), 800)
for _ in range(3):
    rec._recognize_raw(x, lens, want_probs=False)
os.environ['KB_LSTM_DBG'] = sys.argv[1] if len(sys.argv) > 1 else '1'
rec._recognize_raw(x, lens, want_probs=False)
os.environ.pop('KB_LSTM_DBG')
m.set_timing(True)
acc = {}
for _ in range(10):
    rec._recognize_raw(x, lens, want_probs=False)
    for k, v in m.last_timing():
        acc[k] = acc.get(k, 0.0) + v / 10
print('stages ms:', {k: round(v, 4) for k, v in acc.items()}, 'sum', round(sum(acc.values()), 4), file=sys.stderr)
# wall clock of strictly serial calls without the stage timers (host turn-around included)
import time
m.set_timing(False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    rec._recognize_raw(x, lens, want_probs=False)
torch.cuda.synchronize()
print('serial wall ms/call (timing off):', round((time.perf_counter() - t0) * 20, 4), file=sys.stderr)
m.set_timing(True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    rec._recognize_raw(x, lens, want_probs=False)
    m.last_timing()
torch.cuda.synchronize()
print('serial wall ms/call (timing on):', round((time.perf_counter() - t0) * 20, 4), file=sys.stderr)
