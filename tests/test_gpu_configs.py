"""GPU parity on BASELINE.json's own configurations at their stated sizes and on TRAINED weights (through the C ABI):

  * Gallicorpora+_best / all_arabic_scripts (H = 120, 3x13 / 3x9 kernels, 3 x BiLSTM-200, fp16-stored weights) on the reference's
    own test lines: label tuples and strings identical to the reference's, logits <= 1e-3 relative, and no call needed the
    fp32 re-run for activations outside the fp16 operand range;
  * cfg3: kraken's shipped blla.mlmodel on 3 x 1800 x 1350 pages, batch 8;
  * cfg4: two models, widths U{200..2000}, two GPUs when the box has them;
  * cfg5: 64 x 48 x 1200 lines;
  * label exactness of cfg2 over 8 seeds x 64 lines x T in {200, 300, 500}.
"""
import json
import os
import zlib

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden

import fixtures as fx
import kraken_b200 as kb
import vgsl_oracle as vo
from kraken_b200.rpred import pad_batch

pytestmark = pytest.mark.gpu

REL_TOL = 1e-3
CFG2 = '[1,48,0,1 Cr3,3,32 Mp2,2 Cr3,3,64 Mp2,2 S1(1x0)1,3 Lbx256 O1c200]'


def rel_err(a, b):
    a = torch.as_tensor(a).float().cpu()
    b = torch.as_tensor(b).float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def triples(dec):
    return [[(int(t[0]), int(t[1]), int(t[2])) for t in d] for d in dec]


@pytest.mark.parametrize('name', ['trained_gallicorpora', 'trained_arabic'])
def test_trained_recogniser_lines(name):
    g = load_golden(name)
    codec = json.loads(str(g['codec']))
    m = kb.TorchVGSLModel(vgsl=str(g['spec']), codec=codec, model_type=['recognition'], seg_type=str(g['seg_type']),
                          one_channel_mode=str(g['one_channel_mode']))
    m.load_state_dict(fx.trained_weights(g))
    rec = kb.TorchSeqRecognizer(m, device='cuda:0')
    xs = fx.trained_lines(g)
    worst = 0.0
    # the reference's legacy path: one line per call (kraken/rpred.py:297-301)
    for i, x in enumerate(xs):
        dec = rec.predict_labels(x)
        exp = fx.trained_expected(g, i)
        assert triples(dec) == triples([exp]), (name, i)
        assert np.allclose([t[3] for t in dec[0]], [t[3] for t in exp], atol=1e-3)
        assert rec.predict_string(x) == [str(g[f'raw::{i}'])]
        if f'logits::{i}' in g:
            logits, _ = m.nn(x.cuda())
            e = rel_err(logits, g[f'logits::{i}'])
            worst = max(worst, e)
            assert e <= REL_TOL, (name, i, e)
    # the new API's batches (kraken/lib/vgsl/rpred.py:126-131): arrival order, zero right-padding, seq_lens; oracle on the same batch
    om = vo.OracleModel(str(g['spec']), fx.trained_weights(g))
    for k in range(0, len(xs), 8):
        seqs, lens = pad_batch([x[0] for x in xs[k:k + 8]])
        ref_logits, _, ref_olens, ref_dec = vo.rec_predict(om, seqs, lens)
        logits, olens = m.nn(seqs.cuda(), lens)
        e = rel_err(logits, ref_logits)
        worst = max(worst, e)
        assert e <= REL_TOL, (name, k, e)
        assert olens.tolist() == ref_olens.tolist()
        dec = rec.predict_labels(seqs, lens)
        assert triples(dec) == triples(ref_dec), (name, k)
    print(f'[{name}] worst logits rel err {worst:.3e}; range_fallback_count {m.range_fallback_count}')
    assert m.range_fallback_count == 0


def _blla_real():
    g = load_golden('trained_blla')
    x = fx.blla_page_tensor(os.path.join(GOLDEN, 'page_input.webp'))
    m = kb.TorchVGSLModel(vgsl=str(g['spec']), model_type=['segmentation'])
    m.load_state_dict(fx.trained_weights(g))
    m.to('cuda:0')
    return g, x, m


def test_cfg3_real_blla_weights_one_page():
    """kraken/blla.mlmodel on one 3 x 1800 x 1350 page against the reference's own output."""
    from kraken_b200.blla import segmentation_heatmap
    g, x, m = _blla_real()
    if zlib.crc32(x.numpy().tobytes()) != int(g['x_crc']):
        pytest.skip('PIL resize of the page differs from the build container (different Pillow build)')
    logits, _ = m.nn(x.cuda())
    assert tuple(logits.shape) == (1, 4, 450, 338)
    e = rel_err(logits, g['logits'])
    print(f'[cfg3 real weights] logits rel err {e:.3e}; range_fallback_count {m.range_fallback_count}')
    assert e <= REL_TOL, e
    hm = segmentation_heatmap(m, x.cuda(), (1800, 1350))
    assert float((hm[:, :, 3::7, 2::7].cpu() - torch.from_numpy(g['heatmap_sub'].astype(np.float32))).abs().max()) < 2e-3
    assert m.range_fallback_count == 0


def test_cfg3_full_batch_of_eight_pages():
    """BASELINE cfg3 at its stated size: 8 x 3 x 1800 x 1350, real weights.  The reference never batches pages (spred.py:268), so
    every page of the batch must equal its own single-page result; two of them are checked against the oracle."""
    from kraken_b200.blla import segmentation_heatmap
    g, x, m = _blla_real()
    om = vo.OracleModel(str(g['spec']), fx.trained_weights(g))
    pages = [x[0], x[0].flip(2), x[0].flip(1), x[0].roll(97, 2), 1.0 - x[0], x[0].roll(211, 1), x[0].flip(1).flip(2), (x[0] * 0.5 + 0.25)]
    batch = torch.stack(pages).contiguous()
    out, _ = m.nn(batch.cuda())
    assert tuple(out.shape) == (8, 4, 450, 338)
    for i in (1, 6):
        ref, _ = om.forward(batch[i:i + 1], None)
        e = rel_err(out[i:i + 1], ref)
        assert e <= REL_TOL, (i, e)
    for i in (0, 3, 7):
        single, _ = m.nn(batch[i:i + 1].cuda())
        assert rel_err(out[i:i + 1], single) <= 1e-5, i
    hm = segmentation_heatmap(m, batch.cuda(), (1800, 1350))
    assert tuple(hm.shape) == (8, 4, 1800, 1350)
    _, ohm = vo.seg_heatmap(om, batch[6:7], (1800, 1350))
    assert float((hm[6:7].cpu() - ohm).abs().max()) < 1e-3
    assert m.range_fallback_count == 0


def test_cfg4_two_models_mixed_widths():
    """BASELINE cfg4: two recognisers (cfg2 spec, seeds 0 / 1), tags alternating, widths U{200..2000} (seed 1), batches per model in
    arrival order; model B lives on the second GPU when there is one."""
    from collections import defaultdict
    from kraken_b200.rpred import mm_recognize_lines
    ndev = torch.cuda.device_count()
    oms, recs = {}, {}
    for k, (tag, seed) in enumerate((('latin', 0), ('arabic', 1))):
        om = vo.OracleModel(CFG2)
        w = om.init_like_reference(seed)
        m = kb.TorchVGSLModel(vgsl=CFG2)
        m.load_state_dict(w)
        oms[tag], recs[tag] = om, kb.TorchSeqRecognizer(m, device=f'cuda:{k % max(ndev, 1)}')
    g = torch.Generator().manual_seed(1)
    widths = torch.randint(200, 2001, (32,), generator=g).tolist()
    widths[3], widths[4] = 2000, 200
    lines = [torch.rand(1, 48, w, generator=g) for w in widths]
    tags = ['latin', 'arabic'] * 16
    nets = defaultdict(lambda: recs['latin'])
    nets.update(recs)
    out = mm_recognize_lines(nets, lines, tags, batch_size=8)
    for tag in ('latin', 'arabic'):
        idxs = [i for i, t in enumerate(tags) if t == tag]
        for k in range(0, len(idxs), 8):
            chunk = idxs[k:k + 8]
            seqs, lens = pad_batch([lines[i] for i in chunk])
            _, _, _, ref = vo.rec_predict(oms[tag], seqs, lens)
            assert triples([out[i] for i in chunk]) == triples(ref), (tag, k)
    print(f'[cfg4] devices used: {sorted({r.nn._device for r in recs.values()})}')


def test_cfg5_batch_of_1200_wide_lines():
    """BASELINE cfg5's unit of work: 64 x 48 x 1200 lines (T = 300), labels bit-exact, logits <= 1e-3."""
    om = vo.OracleModel(CFG2)
    wts = om.init_like_reference(2)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(64, 1, 48, 1200, generator=g)
    lens = torch.full((64,), 1200, dtype=torch.long)
    ref_logits, _, ref_olens, ref_dec = vo.rec_predict(om, x, lens)
    m = kb.TorchVGSLModel(vgsl=CFG2)
    m.load_state_dict(wts)
    rec = kb.TorchSeqRecognizer(m, device='cuda:0')
    logits, olens = m.nn(x.cuda(), lens)
    assert rel_err(logits, ref_logits) <= REL_TOL
    assert olens.tolist() == ref_olens.tolist() == [300] * 64
    assert triples(rec.predict_labels(x, lens)) == triples(ref_dec)


@pytest.mark.parametrize('w', [800, 1200, 2000])
def test_cfg2_label_exactness_over_seeds(w):
    """8 seeds x 64 lines at T = 200 / 300 / 500: the approximate SFU gates and the split-fp16 operands never flip an arg-max that the
    reference decides by more than its own rounding noise.  Random-init models produce near-uniform logits; over 1536 lines the
    smallest top-2 gap of the fp32 reference is ONE ulp (1.2e-7), where the reference's own arg-max depends on the summation order of
    its BLAS.  Such time steps (gap <= 1e-5 relative, 5x the engine's measured logit error and 100x below the 1e-3 contract) may go
    either way; every other time step, and every line without such a step, must match bit for bit."""
    real_flips, tie_steps, tie_lines, lines_cmp = 0, 0, 0, 0
    min_gap = 1e9
    for seed in range(100, 108):
        om = vo.OracleModel(CFG2)
        wts = om.init_like_reference(seed)
        g = torch.Generator().manual_seed(seed)
        x = torch.rand(64, 1, 48, w, generator=g)
        lens = torch.full((64,), w, dtype=torch.long)
        ref_logits, _, _, ref_dec = vo.rec_predict(om, x, lens)
        rl = ref_logits.squeeze(2)                                   # (N, C, T)
        top2 = rl.topk(2, dim=1).values
        gap = top2[:, 0] - top2[:, 1]                                # (N, T)
        min_gap = min(min_gap, float(gap.min()))
        tol = 1e-5 * float(rl.abs().max())
        m = kb.TorchVGSLModel(vgsl=CFG2)
        m.load_state_dict(wts)
        rec = kb.TorchSeqRecognizer(m, device='cuda:0')
        logits, _ = m.nn(x.cuda(), lens)
        gl = logits.squeeze(2).cpu()
        assert rel_err(gl, rl) <= REL_TOL
        differ = gl.argmax(1) != rl.argmax(1)                        # (N, T)
        real_flips += int((differ & (gap > tol)).sum())
        tie_steps += int((differ & (gap <= tol)).sum())
        dec = triples(rec.predict_labels(x.cuda(), lens))
        ref = triples(ref_dec)
        own = triples(vo.greedy_decode(gl.softmax(1), torch.full((64,), gl.shape[-1])))
        for i in range(64):
            if bool(differ[i].any()):
                tie_lines += 1
                assert dec[i] == own[i]                               # the decoder is consistent with the engine's own logits
            else:
                lines_cmp += 1
                assert dec[i] == ref[i], (seed, i)
    print(f'[seeds] W={w}: smallest top-2 logit gap of the reference {min_gap:.3e}; lines identical to the reference {lines_cmp}/512; '
          f'time steps decided differently at a near-tie {tie_steps} (in {tie_lines} lines); beyond the tie tolerance {real_flips}')
    assert real_flips == 0
    assert tie_lines <= 2
