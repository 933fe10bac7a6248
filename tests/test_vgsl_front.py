"""VGSL front end of the engine (C++ parser behind kb_model_create) against the reference's grammar:
named specs, static shapes, state-dict keys, seq_len arithmetic, error behaviour
(reference tests/test_vgsl.py:43-82; kraken/lib/vgsl/model.py:109-243,570-902)."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

import kraken_b200 as kb
import vgsl_oracle as vo

CFG2 = '[1,48,0,1 Cr3,3,32 Mp2,2 Cr3,3,64 Mp2,2 S1(1x0)1,3 Lbx256 O1c200]'
SPECS = [
    CFG2,
    '[1,30,0,1 Cr3,3,32,2,2 Gn32 Cr3,3,64,2,2 Gn32 S1(1x0)1,3 O1c16]',
    '[1,120,0,1 Cr3,13,32 Do0.1,2 Mp2,2 Cr3,13,32 Do0.1,2 Mp2,2 Cr3,9,64 Do0.1,2 Mp2,2 Cr3,9,64 Do0.1,2 S1(1x0)1,3 Lbx200 Do0.1,2 Lbx200 Do0.1,2 Lbx200 Do O1c80]',
    '[1,1800,0,3 Cr7,7,64,2,2 Gn32 Cr3,3,128,2,2 Gn32 Cr3,3,128 Gn32 Cr3,3,256 Gn32 Cr3,3,256 Gn32 Lbx32 Lby32 Cr1,1,32 Gn32 Lby32 Lbx32 O2l4]',
    '[1,48,0,1 Cr3,3,16 Mp2,2 ([Cr3,3,8 Ct1,1,8] I) S1(1x0)1,3 Lfx16 Lrx8 O1ca10]',
    '[1,48,0,1 Cr3,3,16 ([Cr3,3,8 Ct1,1,8] [Cr3,3,4] I) [Mp2,2] O2l3]',
    '[1,32,0,1 Cr3,3,8 Mp2,2xyz A3,4 Lfys16 Lbx8 O1s7]',
    '[1,24,0,3 Clr5,3,8,1,2 Cm3,3,6 Cl3,5,4,1,1,2,2 Mp3,3,2,2 Cs1,1,5 Lby6 O2s3]',
    '[1,1,0,48 Lbx20 Do O1c59]',
    '[1,48,0,1 Cr{conv_a}3,3,8 Mp{pool}2,2 S{fold}1(1x0)1,3 Lbx{rnn}8 O{out}1c5]',
    '[1,48,0,1 Cr{C_0}3,3,32 Do.{Do_1}1,2 Mp{Mp_2}2,2 S{S_3}1(1x0)1,3 Gbx{G_4}16 O{O_5}1c9]',
]


@pytest.mark.parametrize('spec', SPECS)
def test_named_spec_shapes_and_keys_match_oracle_grammar(spec):
    m = kb.TorchVGSLModel(vgsl=spec)
    om = vo.OracleModel(spec)                     # pinned bit-identically to the reference (tests/test_oracle.py)
    assert '[' + ' '.join(m.named_spec) + ']' == om.named_spec
    assert m.user_metadata['vgsl'] == om.named_spec
    assert tuple(m.input) == tuple(om.input)
    assert tuple(m.output) == tuple(om.output)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v) for k, v in om.param_shapes().items()}
    assert list(m.state_dict().keys()) == list(om.param_shapes().keys())
    # re-parsing the named spec is a fixed point (reference tests/test_vgsl.py:18-41 round trip)
    m2 = kb.TorchVGSLModel(vgsl=m.user_metadata['vgsl'])
    if 'Do.{' not in spec:
        assert m2.user_metadata['vgsl'] == m.user_metadata['vgsl']


def test_append_matches_reference_named_spec():
    # reference tests/test_vgsl.py:43-49
    m = kb.TorchVGSLModel(vgsl='[1,1,0,48 Lbx{foo}64 Do O1c59]')
    m.append(1, '[Cr{bar}2,4,2,2 Gn{baz}2]')
    assert m.user_metadata['vgsl'] == '[1,1,0,48 Lbx{foo}64 Cr{bar}2,4,2,2 Gn{baz}2]'


def test_resize_output():
    # reference tests/test_vgsl.py:51-65
    m = kb.TorchVGSLModel(vgsl='[1,1,0,48 Lbx10 Do O1c57]')
    m.resize_output(80)
    assert m.output[1] == 80 and m.state_dict()['nn.O_2.lin.weight'].shape == (80, 20)
    m = kb.TorchVGSLModel(vgsl='[1,1,0,48 Lbx10 Do O1c57]')
    w = m.state_dict()['nn.O_2.lin.weight']
    m.resize_output(80, [2, 3])
    w2 = m.state_dict()['nn.O_2.lin.weight']
    assert w2.shape == (80, 20)
    keep = [i for i in range(57) if i not in (2, 3)]
    assert np.array_equal(w2[:55].numpy(), w[keep].numpy())


@pytest.mark.parametrize('spec,exc', [
    ('[1,1,0,48 Lbx10 Do O0c57]', ValueError),                       # categorical output
    ('[1,48,0,1 Cr3,3,8 O2c5]', ValueError),                         # CTC on heat map
    ('[1,48,0,1 (Cr3,3,8 Mp2,2)]', ValueError),                      # unequal parallel (tests/test_vgsl.py:77-82)
    ('[1,48,0,1 [Cr3,3,8 Mp2,2]', ValueError),                       # unbalanced
    ('[1,48,0,1 Cr3,3,8 Xq7]', ValueError),
    ('1,48,0,1 Cr3,3,8', ValueError),
    ('[a,b Cr3,3,8]', ValueError),
    ('[1,48,0,1 S2(3x0)1,3]', ValueError),                           # neither high nor low is the source dim
    ('[1,48,0,1 A7,3]', ValueError),
    ('[1,48,0,1 Cr3,3,7 Gn2]', ValueError),
    ('[1,48,0,1 CTr3,3,8]', NotImplementedError),                    # transposed conv: parsed, not executable
])
def test_spec_errors(spec, exc):
    with pytest.raises(exc):
        kb.TorchVGSLModel(vgsl=spec)


def test_missing_spec():
    with pytest.raises(ValueError):
        kb.TorchVGSLModel()


@pytest.mark.parametrize('name', golden_names())
def test_runtime_dims_and_lens_match_reference_outputs(name):
    """kb_model_infer_dims / kb_model_infer_lens (pure host arithmetic) against what the reference produced."""
    g = load_golden(name)
    m = kb.TorchVGSLModel(vgsl=str(g['spec']))
    assert '[' + ' '.join(m.named_spec) + ']' == str(g['named_spec'])
    n, c, h, w = g['x'].shape
    assert m.infer_dims(n, h, w) == tuple(g['logits'].shape)
    if 'lens' in g:
        assert m.infer_lens(h, w, g['lens']).tolist() == g['olens'].tolist()


def test_metadata_properties():
    m = kb.TorchVGSLModel(vgsl=CFG2, model_type=['recognition'], seg_type='bbox', one_channel_mode='1')
    assert m.model_type == ['recognition'] and m.seg_type == 'bbox' and m.one_channel_mode == '1'
    with pytest.raises(ValueError):
        m.one_channel_mode = 'x'
    with pytest.raises(ValueError):
        m.seg_type = 'lines'
    with pytest.raises(ValueError):
        m.model_type = 'alignment'
    assert m.use_legacy_polygons is True
    with pytest.raises(RuntimeError):
        m.load_state_dict({'nn.C_0.co.weight': np.zeros((32, 1, 3, 3), np.float32)})
    with pytest.raises(RuntimeError):
        sd = m.state_dict()
        sd['nn.C_0.co.weight'] = sd['nn.C_0.co.weight'][:5]
        m.load_state_dict(sd)
