"""Stand-alone GPU diagnostic: cfg2 per-stage times under an environment switch given on the command line (values 1 and 0 alternating)."""
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
outs = {}
VAR = sys.argv[1]
VALS = sys.argv[2:4] if len(sys.argv) >= 4 else ['1', '0']
for mode in (VALS[0], VALS[1], VALS[0], VALS[1]):
    os.environ[VAR] = mode
    outs[mode], _ = m.nn(x, lens)
    for _ in range(3):
        rec._recognize_raw(x, lens, want_probs=False)
    m.set_timing(True)
    acc = {}
    for _ in range(20):
        rec._recognize_raw(x, lens, want_probs=False)
        for k, v in m.last_timing():
            acc[k] = acc.get(k, 0.0) + v / 20
    m.set_timing(False)
    print(f'{VAR}={mode}:', {k: round(v, 4) for k, v in acc.items()}, 'sum', round(sum(acc.values()), 4), file=sys.stderr)
print('logits identical for both values:', bool(torch.equal(outs[VALS[0]], outs[VALS[1]])), 'max rel diff', float((outs[VALS[0]] - outs[VALS[1]]).abs().max() / outs[VALS[1]].abs().max()), file=sys.stderr)
