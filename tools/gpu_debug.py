"""Stand-alone GPU diagnostic (not a pytest module): per-layer error of the engine vs the oracle for every golden
case, printed as a table.  Usage under gpurun:  python tools/gpu_debug.py > gpurun_out/debug.log 2>&1"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np
import torch

os.environ.setdefault('KB_FUSE', '0')       # per-layer taps need every layer materialised
import __graft_entry__ as ge

ge.build()
import kraken_b200 as kb
import vgsl_oracle as vo
from conftest import golden_names, load_golden

print(torch.cuda.get_device_name(0), 'devices', kb.device_count(), flush=True)
for name in golden_names(('cfg1', 'cfg2', 'rec_', 'seg_', 'misc_')):
    try:
        g = load_golden(name)
        spec = str(g['spec'])
        om = vo.OracleModel(spec)
        if any(k.startswith('w::') for k in g):
            w = {k[3:]: g[k] for k in g if k.startswith('w::')}
            om.load(w)
        else:
            w = om.init_like_reference(int(g['seed']))
        x = torch.from_numpy(g['x'])
        lens = torch.from_numpy(g['lens']) if 'lens' in g else None
        taps = {}
        ref, rl = om.forward(x, lens, taps)
        m = kb.TorchVGSLModel(vgsl=spec)
        m.load_state_dict(w)
        m.to('cuda:0')
        t0 = time.time()
        out, ol = m.nn(x.cuda(), lens)
        torch.cuda.synchronize()
        print(f'== {name}: out {tuple(out.shape)} {time.time() - t0:.3f}s', flush=True)
        for ln, t in taps.items():
            try:
                o = m.nn.layer_output(ln)
                if tuple(o.shape) != tuple(t.shape):
                    print(f'   {ln:10s} SHAPE {tuple(o.shape)} vs {tuple(t.shape)}')
                    continue
                err = float((o - t).abs().max())
                print(f'   {ln:10s} {str(tuple(t.shape)):22s} maxabs {err:.3e}  rel {err / max(float(t.abs().max()), 1e-30):.3e}')
            except Exception as e:
                print(f'   {ln:10s} ERROR {e}')
        print(f'   final rel {float((out.cpu() - ref).abs().max() / ref.abs().max()):.3e} vs golden '
              f'{float((out.cpu() - torch.from_numpy(g["logits"])).abs().max() / np.abs(g["logits"]).max()):.3e}; lens {None if ol is None else ol.tolist()} ref {None if rl is None else rl.tolist()}', flush=True)
    except Exception:
        print(f'== {name}: EXCEPTION')
        traceback.print_exc()
        sys.stdout.flush()
