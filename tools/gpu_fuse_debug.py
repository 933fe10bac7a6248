"""GPU diagnostic for the fused groups: tap error of each group output vs the oracle under the fusion toggles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch
os.environ['KB_KEEP_FP32'] = '1'
import __graft_entry__ as ge
ge.build()
import kraken_b200 as kb
import vgsl_oracle as vo

def run(spec, h, w, n=2):
    om = vo.OracleModel(spec); wts = om.init_like_reference(41)
    g = torch.Generator().manual_seed(41)
    x = torch.rand(n, 1, h, w, generator=g)
    taps = {}
    ref, _ = om.forward(x, None, taps)
    m = kb.TorchVGSLModel(vgsl=spec); m.load_state_dict(wts); m.to('cuda:0')
    for fuse in ('0', '1', '2', '3'):
        os.environ['KB_FUSE'] = fuse; bo = '-'
        out, _ = m.nn(x.cuda())
        line = f'  fuse={fuse} baseoff={bo}: final rel {float((out.cpu()-ref).abs().max()/ref.abs().max()):.2e} |'
        for name, t in taps.items():
            try:
                o = m.nn.layer_output(name)
                err = float((o - t).abs().max() / t.abs().max().clamp_min(1e-30))
                line += f' {name}:{err:.1e}'
                if err > 1e-3 and o.dim() == 4 and fuse != '1':
                    d = (o - t).abs()
                    # where is it wrong? per-column (w) and per-channel profile of the first bad tap
                    bad_w = (d.amax(dim=(0, 1, 2)) > 1e-3).nonzero().flatten().tolist()
                    bad_c = (d.amax(dim=(0, 2, 3)) > 1e-3).nonzero().flatten().tolist()
                    line += f'[bad w {bad_w[:6]}..{len(bad_w)}/{o.shape[3]} bad c {bad_c[:4]}..{len(bad_c)}/{o.shape[1]}]'
            except Exception as e:
                pass
        print(line, flush=True)

print(torch.cuda.get_device_name(0))
for spec, h, w in [('[1,48,0,1 Cr3,3,32 Mp2,2 Cr3,3,64 Mp2,2 S1(1x0)1,3 Lbx32 O1c20]', 48, 512),
                   ('[1,8,0,1 Cr3,3,32 Mp2,2 Cr1,3,32 O1c20]', 8, 256),
                   ('[1,8,0,1 Cr3,3,32 Mp2,2 Cr3,1,32 O1c20]', 8, 256)]:
    print(spec, h, w)
    run(spec, h, w)
