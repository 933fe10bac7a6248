"""Stand-alone GPU workload for profilers (not a pytest module): `steps` strictly serial cfg2 batches (64 x 48 x 800, random-init
weights from the engine's own initialiser) through kb_recognize.  Usage under gpurun:
  ncu --set full --clock-control none --import-source on -s 24 -c 8 -o gpurun_out/r02_prof python tools/one_step.py 4"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import kraken_b200 as kb

CFG2 = '[1,48,0,1 Cr3,3,32 Mp2,2 Cr3,3,64 Mp2,2 S1(1x0)1,3 Lbx256 O1c200]'
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
m = kb.TorchVGSLModel(vgsl=CFG2)
m.init_weights()
rec = kb.TorchSeqRecognizer(m, device='cuda:0')
x = torch.rand(64, 1, 48, 800).cuda()
lens = torch.full((64,), 800)
for _ in range(steps):
    r = rec._recognize_raw(x, lens, want_probs=False)
torch.cuda.synchronize()
print('decoded labels in the last batch:', int(r['counts'].sum()), 'launches per step:', m.launch_count // steps)
