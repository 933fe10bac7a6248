"""The oracle against the fixtures of the TRAINED models (tests/golden/trained_*.npz, generated from the unmodified reference by
`oracle/make_golden.py --trained`): Gallicorpora+_best (H = 120, 3x13 / 3x9 kernels, 3 x BiLSTM-200) on the 29 bbox lines of the
reference's input.webp, all_arabic_scripts on arabic.webp, and kraken's shipped blla.mlmodel on one 1800 x 1350 page (cfg3)."""
import json
import zlib
from difflib import SequenceMatcher

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden

import fixtures as fx
import vgsl_oracle as vo
from kraken_b200.codec import PytorchCodec

TOL = 2e-5


@pytest.mark.parametrize('name,lines', [('trained_gallicorpora', (0, 3, 7, 20)), ('trained_arabic', (1, 5))])
def test_oracle_reproduces_trained_recognisers(name, lines):
    g = load_golden(name)
    om = vo.OracleModel(str(g['spec']), fx.trained_weights(g))
    codec = PytorchCodec(json.loads(str(g['codec'])))
    xs = fx.trained_lines(g)
    assert len(xs) == int(g['n_lines'])
    for i in lines:
        logits, _, _, dec = vo.rec_predict(om, xs[i])
        if f'logits::{i}' in g:
            ref = torch.from_numpy(g[f'logits::{i}'])
            assert float((logits - ref).abs().max()) <= TOL * max(1.0, float(ref.abs().max()))
        exp = fx.trained_expected(g, i)
        assert [t[:3] for t in dec[0]] == [t[:3] for t in exp]
        assert np.allclose([t[3] for t in dec[0]], [t[3] for t in exp], atol=1e-4)
        raw = ''.join(c for c, *_ in codec.decode(dec[0]))
        assert raw == str(g[f'raw::{i}'])
        if f'expected::{i}' in g:
            # criterion of the reference's tests/test_tasks.py:117-130 against box_rec.pkl (the reference's record additionally went
            # through BiDi reordering, which is the identity for this left-to-right text)
            assert SequenceMatcher(isjunk=None, a=str(g[f'pred::{i}']), b=str(g[f'expected::{i}'])).ratio() > 0.9


def test_oracle_reproduces_real_blla_page():
    g = load_golden('trained_blla')
    import os
    x = fx.blla_page_tensor(os.path.join(GOLDEN, 'page_input.webp'))
    if zlib.crc32(x.numpy().tobytes()) != int(g['x_crc']):
        pytest.skip('PIL resize of the page differs from the build container (different Pillow build)')
    om = vo.OracleModel(str(g['spec']), fx.trained_weights(g))
    logits, hm = vo.seg_heatmap(om, x, (1800, 1350))
    ref = torch.from_numpy(g['logits'])
    assert float((logits - ref).abs().max()) <= TOL * float(ref.abs().max())
    assert float((hm[:, :, 3::7, 2::7] - torch.from_numpy(g['heatmap_sub'].astype(np.float32))).abs().max()) < 2e-3
