"""Debug helper (not a test): run one small-hidden LSTM spec through the engine and compare with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kraken_b200 as kb
from oracle import vgsl_oracle as vo

spec = sys.argv[1] if len(sys.argv) > 1 else '[1,16,0,1 Cr3,3,32 Mp2,2 S1(1x0)1,3 Lbx32 O1c30]'
om = vo.OracleModel(spec)
wts = om.init_like_reference(77)
g = torch.Generator().manual_seed(77)
x = torch.rand(70, om.input[1], 16, 120, generator=g)
ref, _ = om.forward(x, None)
m = kb.TorchVGSLModel(vgsl=spec)
m.load_state_dict(wts)
m.to('cuda:0')
out, _ = m.nn(x.cuda(), None)
torch.cuda.synchronize()
err = (out.cpu() - ref).abs().max().item() / ref.abs().max().item()
print('rel err', err)
