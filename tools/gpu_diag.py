"""Stand-alone GPU diagnostic (not a pytest module): where does an error against the oracle come from?

For the trained Gallicorpora+ recogniser (line 0 of the golden fixture) and a small cfg2 batch the final logits are compared with the
oracle under every run-time switch that moves one kind of layer back to the fp32 CUDA-core kernels, then layer by
layer with the fused groups off.  Usage under gpurun:  timeout 300 python tools/gpu_diag.py > gpurun_out/diag.log 2>&1"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import faulthandler

faulthandler.dump_traceback_later(240, exit=True)
import numpy as np
import torch

import kraken_b200 as kb
import vgsl_oracle as vo
import fixtures as fx
from conftest import load_golden

CFG2 = '[1,48,0,1 Cr3,3,32 Mp2,2 Cr3,3,64 Mp2,2 S1(1x0)1,3 Lbx256 O1c200]'
SWITCHES = [{}, {'KB_LSTM_TC': '0'}, {'KB_GEMM': 'ffma'}, {'KB_FUSE': '0'}, {'KB_FUSE': '1'}, {'KB_FUSE': '2'}, {'KB_LSTM_GENERIC': '1'}]


def rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).abs().max() / b.abs().max())


def under(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def case(name, spec, w, x, lens):
    om = vo.OracleModel(spec, w) if w is not None else vo.OracleModel(spec)
    if w is None:
        w = om.init_like_reference(0)
    taps = {}
    ref, _ = om.forward(x, lens, taps)
    for env in SWITCHES:
        def run():
            m = kb.TorchVGSLModel(vgsl=spec)
            m.load_state_dict(w)
            m.to('cuda:0')
            out, _ = m.nn(x.cuda(), lens)
            torch.cuda.synchronize()
            return out.cpu(), m
        t0 = time.time()
        out, m = under(env, run)
        print(f'{name:14s} {str(env):28s} final rel {rel(out, ref):.3e}   ({time.time() - t0:.2f}s, fallbacks {m.range_fallback_count})', flush=True)
    def layers():
        m = kb.TorchVGSLModel(vgsl=spec)
        m.load_state_dict(w)
        m.to('cuda:0')
        m.nn(x.cuda(), lens)
        for ln, t in taps.items():
            try:
                o = m.nn.layer_output(ln)
                print(f'   {ln:10s} {str(tuple(t.shape)):24s} rel {rel(o, t):.3e}')
            except Exception as e:
                print(f'   {ln:10s} ERROR {e}')
    print(f'{name}: layer by layer with KB_FUSE=0 (tensor-core layers on)')
    under({'KB_FUSE': '0'}, layers)
    print(f'{name}: layer by layer with KB_FUSE=0 KB_GEMM=ffma KB_LSTM_TC=0 (all fp32 CUDA-core)')
    under({'KB_FUSE': '0', 'KB_GEMM': 'ffma', 'KB_LSTM_TC': '0'}, layers)


print(torch.cuda.get_device_name(0), flush=True)
g = load_golden('trained_gallicorpora')
xs = fx.trained_lines(g)
case('gallicorpora/0', str(g['spec']), fx.trained_weights(g), xs[0], None)
gen = torch.Generator().manual_seed(5)
x2 = torch.rand(16, 1, 48, 400, generator=gen)
case('cfg2 16x400', CFG2, None, x2, torch.full((16,), 400, dtype=torch.long))
