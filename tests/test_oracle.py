"""The oracle against the committed golden fixtures (generated from the unmodified reference by
oracle/make_golden.py) and - when the reference tree is present (build container only) - against the reference
itself, bit for bit."""
import os

import numpy as np
import pytest
import torch

from conftest import dec_from_golden, golden_names, load_golden

import vgsl_oracle as vo

TOL = 2e-5      # same ATen kernels, but the GPU box may have a different CPU (different oneDNN/MKL code paths)


def model_for(g):
    om = vo.OracleModel(str(g['spec']))
    if any(k.startswith('w::') for k in g):
        om.load({k[3:]: g[k] for k in g if k.startswith('w::')})
    else:
        om.init_like_reference(int(g['seed']))
    return om


@pytest.mark.parametrize('name', golden_names(('cfg1', 'cfg2', 'rec_', 'seg_', 'misc_')))
def test_oracle_reproduces_reference_goldens(name):
    g = load_golden(name)
    om = model_for(g)
    x = torch.from_numpy(g['x'])
    lens = torch.from_numpy(g['lens']) if 'lens' in g else None
    logits, olens = om.forward(x, lens)
    ref = torch.from_numpy(g['logits'])
    assert logits.shape == ref.shape
    assert float((logits - ref).abs().max()) <= TOL * max(1.0, float(ref.abs().max()))
    if 'olens' in g:
        assert olens.tolist() == g['olens'].tolist()
    if 'dec_count' in g:
        temp = float(g['temperature']) if 'temperature' in g else 1.0
        _, probs, ol, dec = vo.rec_predict(om, x, lens, temp)
        exp = dec_from_golden(g)
        assert [[t[:3] for t in d] for d in dec] == [[t[:3] for t in d] for d in exp]
        for d, e in zip(dec, exp):
            assert np.allclose([t[3] for t in d], [t[3] for t in e], atol=1e-5)
    if 'heatmap' in g:
        _, hm = vo.seg_heatmap(om, x, tuple(g['seg_size'].tolist()))
        assert float((hm - torch.from_numpy(g['heatmap'].astype(np.float32))).abs().max()) < 2e-3


def test_cfg1_golden_strings():
    """reference tests/test_rpred.py:352-358 and :453-462 - exact strings through oracle + codec."""
    import json
    from kraken_b200.codec import PytorchCodec
    for name in ('cfg1_overfit_bbox', 'cfg1_overfit_nobidi'):
        g = load_golden(name)
        om = model_for(g)
        _, _, _, dec = vo.rec_predict(om, torch.from_numpy(g['x']))
        codec = PytorchCodec(json.loads(str(g['codec'])))
        raw = ''.join(c for c, *_ in codec.decode(dec[0]))
        assert raw == str(g['raw_prediction'])
    assert str(load_golden('cfg1_overfit_nobidi')['raw_prediction']) == 'ܕܗܣܐܕ ܪܝ .ܡܡ ܐܠܠ ܗܠ ܐܘܗ ܟܘܗܢ ܡܡ ܐܠ'
    # display-order string of the bbox test is the BiDi reordering of the raw one (BiDi itself is out of scope)
    assert str(load_golden('cfg1_overfit_bbox')['prediction']) == 'ܡ ܘܡ ܗ ܡܕܐ ܐ ܐܐ ܡ ܗܗܐܐܐܕ'


def test_greedy_decode_edge_cases():
    p = torch.zeros(3, 6)
    p[0, :] = 1.0                       # all blank
    assert vo.greedy_decode(p) == [[]]
    p = torch.tensor([[0.1, 0.1, 0.8, 0.8, 0.1, 0.2], [0.9, 0.7, 0.1, 0.1, 0.2, 0.1], [0.0, 0.2, 0.1, 0.1, 0.7, 0.7]])
    assert vo.greedy_decode(p) == [[(1, 0, 1, pytest.approx(0.9)), (2, 4, 5, pytest.approx(0.7))]]
    with pytest.raises(ValueError):
        vo.greedy_decode(torch.rand(2, 3, 4))
    assert vo.greedy_decode(torch.rand(2, 3, 4), torch.tensor([0, 4]))[0] == []


@pytest.mark.skipif(not os.path.isdir('/root/reference/kraken'), reason='reference tree only exists in the build container')
def test_oracle_is_bit_identical_to_reference():
    import refshim
    refshim.install()
    from kraken.lib.ctc_decoder import greedy_decoder
    from kraken.lib.vgsl.model import TorchVGSLModel
    specs = ['[1,48,0,1 Cr3,3,32 Mp2,2 Cr3,3,64 Mp2,2 S1(1x0)1,3 Lbx256 O1c200]',
             '[1,30,0,1 Cr3,3,32,2,2 Gn32 Cr3,3,64,2,2 Gn32 S1(1x0)1,3 O1c16]',
             '[1,64,0,3 Cr7,7,16,2,2 Gn8 Cr3,3,32,2,2 Gn8 Lbx8 Lby8 Cr1,1,8 Gn4 Lby8 Lbx8 O2l4]',
             '[1,48,0,1 Cr3,3,16 Mp2,2 ([Cr3,3,8 Ct1,1,8] I) S1(1x0)1,3 Lfx16 Lrx8 O1ca10]',
             '[1,32,0,1 Cr3,3,8 Mp2,2xyz A3,4 Lfys16 Lbx8 O1s7]',
             '[1,32,0,1 Cr3,3,8 Mp2,2 S1(1x0)1,3 Lbxc12 Lfxc8 O1c9]',          # legacy clstm cells (ones column, no biases)
             '[1,32,0,1 Cr3,3,8 Mp2,2 S1(1x0)1,3 Lbxo12 O1c9]']                 # legacy ocropy cell (peepholes)
    for sp in specs:
        torch.manual_seed(3)
        ref = TorchVGSLModel(vgsl=sp)
        ref.eval()
        with torch.no_grad():                                                   # PeepholeBidiLSTM leaves its parameters uninitialised
            for k, v in ref.state_dict().items():
                if not torch.isfinite(v).all() or v.abs().max() > 1e3 or '_ip_' in k or '_fp_' in k or '_op_' in k or ('o12' in sp and '.layer.' in k):
                    v.copy_(torch.rand(v.shape) * 0.4 - 0.2)
        om = vo.OracleModel(sp, dict(ref.state_dict()))
        assert om.named_spec == ref.user_metadata['vgsl']
        assert tuple(om.output) == tuple(ref.output)
        x = torch.rand(3, om.input[1], om.input[2] or 1, 80)
        for lens in (None, torch.tensor([80, 41, 13])):
            try:
                with torch.inference_mode():
                    ro, rl = ref.nn(x, lens)
            except Exception:
                with pytest.raises(Exception):
                    om.forward(x, lens)
                continue
            oo, ol = om.forward(x, lens)
            assert torch.equal(ro, oo)
            assert (rl is None and ol is None) or rl.tolist() == ol.tolist()
            if ro.shape[2] == 1:
                p = ro.softmax(1).squeeze(2)
                ll = rl if rl is not None else torch.tensor([p.shape[-1]] * 3)
                assert greedy_decoder(p, ll) == vo.greedy_decode(p, ll)
