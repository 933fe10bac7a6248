"""Record assembly: code-point lookup + `_scale_val` position scaling inside the CTC collapse kernel.

CPU: the formula the kernel implements, restated in numpy doubles, is pinned against the reference's own method
(kraken/lib/vgsl/rpred.py:231 `_scale_val`, imported through oracle/refshim.py where the reference is mounted) and against a pure-Python
copy of it everywhere else.  GPU: `kb_recognize_records` against `kb_recognize` + `PytorchCodec.decode` + that formula."""
import os

import numpy as np
import pytest
import torch

import kraken_b200 as kb
from kraken_b200.codec import PytorchCodec

CFG2 = '[1,48,0,1 Cr3,3,32 Mp2,2 Cr3,3,64 Mp2,2 S1(1x0)1,3 Lbx256 O1c200]'


def scale_val_py(val, net_scale, padding, in_scale, min_val, max_val):
    """verbatim arithmetic of rpred.py:231 (Python floats, built-in round)"""
    return int(round(min(max(((val * net_scale) - padding) * in_scale, min_val), max_val - 1)))


def scale_val_np(val, net_scale, padding, in_scale, max_val):
    """what the kernel does: separately rounded double multiply / subtract / multiply, clamp, round-half-even"""
    x = np.float64(val) * np.float64(net_scale)
    x = (x - np.float64(padding)) * np.float64(in_scale)
    x = x if x > 0.0 else np.float64(0.0)
    hi = np.float64(max_val - 1)
    x = x if x < hi else hi
    return int(np.rint(x))


def test_scale_val_restatement_equals_python_round_semantics():
    rng = np.random.default_rng(0)
    for _ in range(20000):
        w = int(rng.integers(40, 3000)); olen = int(rng.integers(1, 800)); pad = int(rng.choice([0, 16])); ow = int(rng.integers(5, 5000))
        if w - 2 * pad <= 0:
            continue
        ns, isc = w / olen, ow / (w - 2 * pad)
        v = int(rng.integers(0, olen + 1))
        assert scale_val_np(v, ns, pad, isc, ow) == scale_val_py(v, ns, pad, isc, 0, ow)
    # exact .5 cases: Python rounds half to even, and so does rint
    assert scale_val_np(1, 2.5, 0, 1.0, 100) == scale_val_py(1, 2.5, 0, 1.0, 0, 100) == 2
    assert scale_val_np(1, 3.5, 0, 1.0, 100) == scale_val_py(1, 3.5, 0, 1.0, 0, 100) == 4


@pytest.mark.skipif(not os.path.isdir('/root/reference/kraken'), reason='reference checkout not mounted (build container only)')
def test_scale_val_copy_is_the_references_method():
    import refshim
    refshim.install()                       # serves stub packages for the reference's missing dependencies
    from kraken.lib.vgsl.rpred import VGSLRecognitionInference as Ref   # noqa
    import inspect
    src = inspect.getsource(Ref._scale_val)
    assert 'int(round(min(max(((val * self.net_scale) - self._inf_config.padding) * self.in_scale, min_val), max_val - 1)))' in src
    # and the method itself on a bare instance (no model needed for the arithmetic)
    import types
    obj = Ref.__new__(Ref)
    obj._inf_config = types.SimpleNamespace(padding=16)
    rng = np.random.default_rng(1)
    for _ in range(3000):
        w = int(rng.integers(40, 3000)); olen = int(rng.integers(1, 800)); ow = int(rng.integers(5, 5000))
        obj.net_scale, obj.in_scale = w / olen, ow / (w - 32)
        v = int(rng.integers(0, olen + 1))
        assert obj._scale_val(v, 0, ow) == scale_val_np(v, obj.net_scale, 16, obj.in_scale, ow)


@pytest.mark.gpu
@pytest.mark.parametrize('u8', [False, True])
def test_gpu_records_equal_decode_plus_scale_val(u8):
    import vgsl_oracle as vo
    om = vo.OracleModel(CFG2)
    wts = om.init_like_reference(8)
    codec = PytorchCodec({chr(0x61 + i) if i < 26 else chr(0x0400 + i): [i + 1] for i in range(150)})     # 1:1, labels 1..150; 151..199 unknown
    m = kb.TorchVGSLModel(vgsl=CFG2, codec=codec)
    m.load_state_dict(wts)
    rec = kb.TorchSeqRecognizer(m, device='cuda:0')
    g = torch.Generator().manual_seed(8)
    n, w, pad = 24, 640, 16
    lens = torch.randint(80, w + 1, (n,), generator=g); lens[0] = w
    x = torch.rand(n, 1, 48, w, generator=g)
    if u8:
        x = (x * 255).to(torch.uint8)
    for i, l in enumerate(lens.tolist()):
        x[i, ..., l:] = 0
    orig = torch.randint(50, 4000, (n,), generator=g).numpy().astype(np.int32)
    inv = np.full(n, 255, np.int16) if u8 else None
    got = rec.recognize_records(x, lens, orig, padding=pad, invert_max=inv)
    raw = rec.recognize_u8(x, lens.numpy().astype(np.int32), inv) if u8 else rec._recognize_raw(x, lens, want_probs=False)
    for i in range(n):
        c = int(raw['counts'][i]); olen = int(raw['olens'][i]); wi = int(lens[i])
        trip = [(int(raw['labels'][i, j]), int(raw['starts'][i, j]), int(raw['ends'][i, j]), float(raw['confs'][i, j])) for j in range(c)]
        dec = codec.decode(trip)                                          # skips the labels outside the codec (strict = False)
        ns, isc = wi / olen, int(orig[i]) / (wi - 2 * pad)
        exp_text = ''.join(d[0] for d in dec)
        exp_pos = [[scale_val_py(d[1], ns, pad, isc, 0, int(orig[i])), scale_val_py(d[2], ns, pad, isc, 0, int(orig[i]))] for d in dec]
        assert got[i][0] == exp_text, i
        assert got[i][1] == exp_pos, i
        assert np.allclose(got[i][2], [d[3] for d in dec], atol=1e-6)
    assert any(len(t[0]) > 0 for t in got)
