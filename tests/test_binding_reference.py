"""The reference-side binding driven by the REAL kraken code paths, in the build container (the reference tree does not
travel to the GPU box; there is no GPU here, so the engine calls are substituted by the oracle - this tests the plumbing of cfg1,
which is a CPU configuration by definition):

  * legacy API: `kraken.rpred.rpred / mm_rpred` (kraken/rpred.py:61-391) driven with the MIRROR object `kraken_b200.TorchSeqRecognizer`
    - attributes touched at rpred.py:99-104,119-124,162-166,226-231,297-301,330-338;
  * new API: the reference's own `TorchVGSLModel.predict(im, segmentation)` (kraken/lib/vgsl/rpred.py:54-229) on a model passed through
    `kraken_b200.accel.accelerate`, including `return_logits` (rpred.py:200,227) and the `decoder` hook (configs/base.py:235);
  * the entry points pyproject.toml registers.
"""
import contextlib
import json
import os
import sys
import warnings

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'kraken')), reason='reference tree only exists in the build container')

import vgsl_oracle as vo  # noqa: E402


def _oracle_blocks(weights, spec, line, lens, temperature=1.0, want_probs=False, rec=None):
    """What kb_recognize returns, computed by the oracle."""
    om = vo.OracleModel(spec, {k: torch.as_tensor(np.asarray(v)) for k, v in weights.items()})
    x = torch.as_tensor(line).float()
    logits, probs, olens, dec = vo.rec_predict(om, x, None if lens is None else torch.as_tensor(lens), temperature)
    n, T = x.shape[0], probs.shape[-1]
    lab = np.zeros((n, T), np.int32); st = np.zeros((n, T), np.int32); en = np.zeros((n, T), np.int32); cf = np.zeros((n, T), np.float32)
    cnt = np.zeros(n, np.int32)
    for i, d in enumerate(dec):
        cnt[i] = len(d)
        for j, (l, s, e, c) in enumerate(d):
            lab[i, j], st[i, j], en[i, j], cf[i, j] = l, s, e, c
    if rec is not None:
        rec.outputs = probs.numpy() if want_probs else type('S', (), {'shape': tuple(probs.shape)})()
    return {'labels': lab, 'starts': st, 'ends': en, 'confs': cf, 'counts': cnt,
            'olens': None if lens is None else np.asarray(olens, np.int32)}


@pytest.fixture(scope='module')
def ref():
    import refshim
    refshim.install()
    warnings.simplefilter('ignore')
    return refshim


def test_legacy_rpred_drives_the_mirror_object(ref, monkeypatch):
    """reference tests/test_rpred.py:352-358 and :453-462 with the mirror recogniser in place of kraken.lib.models.TorchSeqRecognizer."""
    from collections import defaultdict
    from PIL import Image
    from kraken.containers import BBoxLine, Segmentation
    from kraken.rpred import mm_rpred, rpred
    import kraken_b200 as kb
    res = os.path.join(REF, 'tests', 'resources')
    m = kb.TorchVGSLModel.load_model(os.path.join(res, 'overfit.mlmodel'))          # native CoreML reader, host only
    rec = kb.TorchSeqRecognizer(m, device=None)
    assert rec.seg_type == 'bbox' and rec.nn.one_channel_mode == '1' and rec.nn.input == (1, 1, 30, 0) or rec.nn.input[2] == 30
    calls = []

    def fake_raw(line, lens, want_probs, out=None):
        calls.append(tuple(line.shape))
        return _oracle_blocks(m.state_dict(), m.spec, line, lens, rec.temperature, want_probs, rec)
    monkeypatch.setattr(rec, '_recognize_raw', fake_raw)
    im = Image.open(os.path.join(res, '000236.png'))
    seg = Segmentation(type='bbox', imagename='000236.png', lines=[BBoxLine(id='foo', bbox=[0, 0, 2544, 156])],
                       text_direction='horizontal-lr', script_detection=False)
    record = next(rpred(rec, im, seg, True))
    assert record.prediction == 'ܡ ܘܡ ܗ ܡܕܐ ܐ ܐܐ ܡ ܗܗܐܐܐܕ'
    g = load_golden('cfg1_overfit_bbox')
    assert record.cuts == [[list(p) if not isinstance(p, list) else p for p in c] for c in g['cuts'].tolist()] or np.array_equal(np.asarray(record.cuts), g['cuts'])
    assert np.allclose(record.confidences, g['confidences'], atol=1e-5)
    record = next(mm_rpred(defaultdict(lambda: rec), im, seg, bidi_reordering=False))
    assert record.prediction == 'ܕܗܣܐܕ ܪܝ .ܡܡ ܐܠܠ ܗܠ ܐܘܗ ܟܘܗܢ ܡܡ ܐܠ'
    assert calls and all(c[0] == 1 and c[1] == 1 for c in calls)               # the legacy path feeds one line per call


class _Pool:
    def imap_unordered(self, func, iterable):
        for item in iterable:
            yield func(item)

    def terminate(self):
        return None


class _Fabric:
    def init_tensor(self):
        return contextlib.nullcontext()


def _prepared_reference_model(ref, **cfg):
    """The reference model as `prepare_for_inference` leaves it (model.py:491-525), minus Fabric (not installed here)."""
    from kraken.configs import RecognitionInferenceConfig
    from kraken.lib.vgsl.model import TorchVGSLModel
    from kraken_b200.weights import load_coreml
    mf = load_coreml(os.path.join(REF, 'tests', 'resources', 'overfit.mlmodel'))[0]
    m = TorchVGSLModel(vgsl=mf.vgsl, codec=mf.codec)
    m.load_state_dict({k: torch.as_tensor(np.asarray(v)).float() for k, v in mf.weights.items()})
    m.user_metadata.update(mf.metadata)
    m.eval()
    m._inf_config = RecognitionInferenceConfig(num_line_workers=0, **cfg)
    m._line_extraction_pool = _Pool()
    m._fabric = _Fabric()
    m._m_dtype = torch.float32
    return m, mf


def test_accelerate_rebinds_the_new_api(ref, monkeypatch):
    from PIL import Image
    from kraken.containers import BBoxLine, Segmentation
    from kraken_b200 import accel
    im = Image.open(os.path.join(REF, 'tests', 'resources', '000236.png'))
    seg = Segmentation(type='bbox', imagename='000236.png', lines=[BBoxLine(id='foo', bbox=[0, 0, 2544, 156]), BBoxLine(id='bar', bbox=[0, 10, 1200, 150])],
                       text_direction='horizontal-lr', script_detection=False)
    stock, _ = _prepared_reference_model(ref, batch_size=2)
    want = [r for r in stock.predict(im, seg)]
    m, mf = _prepared_reference_model(ref, batch_size=2)
    calls = []

    def fake_recognize(rec, line, lens, want_probs):
        calls.append((tuple(line.shape), None if lens is None else lens.tolist(), want_probs, rec.temperature))
        return _oracle_blocks(rec.nn.state_dict(), rec.nn.spec, line, lens, rec.temperature, want_probs, rec)
    monkeypatch.setattr(accel, '_engine_recognize', fake_recognize)
    out = accel.accelerate(m, device=None)
    assert out is m and type(m).__name__ == 'TorchVGSLModel'                    # same object, same class: isinstance checks keep working
    assert sorted(m.state_dict()) == sorted(stock.state_dict())
    got = [r for r in m.predict(im, seg)]
    assert [r.prediction for r in got] == [r.prediction for r in want]
    assert [r.cuts for r in got] == [r.cuts for r in want]                       # _scale_val positions (rpred.py:231) from the engine's starts/ends
    assert all(np.allclose(a.confidences, b.confidences, atol=1e-5) for a, b in zip(got, want))
    assert calls == [((2, 1, 30, calls[0][0][3]), calls[0][1], False, 1.0)] and len(calls[0][1]) == 2     # ONE batched fused call, no probabilities
    # return_logits: `outputs` must be the (N, C, W) probability tensor the record slices (rpred.py:200,227)
    m._inf_config.return_logits = True
    m._inf_config.temperature = 2.0
    stock._inf_config.return_logits = True
    stock._inf_config.temperature = 2.0
    got = [r for r in m.predict(im, seg)]
    want = [r for r in stock.predict(im, seg)]
    assert calls[-1][2] is True and calls[-1][3] == 2.0
    assert torch.is_tensor(m.outputs) and tuple(m.outputs.shape) == tuple(stock.outputs.shape)
    assert float((m.outputs - stock.outputs).abs().max()) < 1e-6
    assert [r.prediction for r in got] == [r.prediction for r in want]
    # decoder hook: a custom decoder gets the probabilities and the output lengths, its result is what is decoded
    seen = {}

    def my_decoder(outputs, seq_lens=None):
        from kraken.lib.ctc_decoder import greedy_decoder
        seen['shape'] = tuple(outputs.shape)
        return [d[:3] for d in greedy_decoder(outputs, seq_lens)]
    m._inf_config.decoder = my_decoder
    m._inf_config.return_logits = False
    got = [r for r in m.predict(im, seg)]
    assert seen['shape'][0] == 2 and all(len(r.prediction) <= 3 for r in got)


def test_nn_forward_is_rebound_for_legacy_callers(ref, monkeypatch):
    """`model.nn(x, lens)` - what kraken.lib.models.TorchSeqRecognizer.forward (models.py:112) and kraken.blla call - goes to the engine,
    while parameters, indexing and the state dict of the module stay in place."""
    from kraken_b200 import accel
    m, _ = _prepared_reference_model(ref)
    seen = []

    def fake_forward(net, x, seq_lens):
        seen.append(tuple(x.shape))
        om = vo.OracleModel(net.spec, net.state_dict())
        return om.forward(x, seq_lens)
    monkeypatch.setattr(accel, '_engine_forward', fake_forward)
    x = torch.rand(2, 1, 30, 80)
    with torch.inference_mode():
        ref_out, ref_l = m.nn(x, torch.tensor([80, 40]))
    accel.accelerate(m, device=None)
    out, ol = m.nn(x, torch.tensor([80, 40]))
    assert seen == [(2, 1, 30, 80)]
    assert torch.equal(out, ref_out) and ol.tolist() == ref_l.tolist()
    assert len(list(m.parameters())) > 0 and type(m.nn[-1]).__name__ == 'LinSoftmax'


def test_entry_points_are_declared_and_resolve(ref):
    import tomllib
    with open(os.path.join(ROOT, 'pyproject.toml'), 'rb') as fp:
        pp = tomllib.load(fp)
    eps = pp['project']['entry-points']
    assert eps['kraken.loaders'] == {'b200': 'kraken_b200.accel:load_accelerated'}
    assert eps['kraken.models'] == {'TorchVGSLModelB200': 'kraken_b200.accel:TorchVGSLModelB200'}
    ref_pp = tomllib.load(open(os.path.join(REF, 'pyproject.toml'), 'rb'))['project']['entry-points']
    assert not set(eps['kraken.models']) & set(ref_pp['kraken.models'])           # names must be unique (kraken/models/utils.py:20-23)
    assert not set(eps['kraken.loaders']) & set(ref_pp['kraken.loaders'])
    from kraken.lib.vgsl.model import TorchVGSLModel
    import kraken_b200.accel as accel
    cls = accel.TorchVGSLModelB200
    assert issubclass(cls, TorchVGSLModel)
    if int(__import__('kraken_b200').lib.kb_device_count()) == 0:
        with pytest.raises(ValueError):                                          # "not mine": load_models falls through to the stock loaders
            accel.load_accelerated(os.path.join(REF, 'tests', 'resources', 'overfit.mlmodel'))
