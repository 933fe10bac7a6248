#!/usr/bin/env python
"""
oracle/make_golden.py - generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/refshim.py) on the CPU in fp32.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference tree does not exist on the
GPU box):      python oracle/make_golden.py

Cases
  cfg1_overfit_bbox   reference tests/test_rpred.py:352-358  rpred(model, 000236.png, bbox seg, pad=True)
                      -> exact string; the captured network input, the model weights
                      (tests/resources/overfit.mlmodel, converted), logits and label tuples are stored.
  cfg1_overfit_nobidi reference tests/test_rpred.py:453-462  mm_rpred(..., bidi_reordering=False), pad=16
  cfg2_small          BASELINE cfg2 spec, seeded weights, 5 ragged lines (incl. width 1-column edge cases)
  rec_default_small   kraken's default recogniser spec (configs/vgsl.py:102) with H=120, 3xBiLSTM-200
  seg_blla_small      the blla.mlmodel architecture on two small 3-channel pages + upsample/sigmoid
  misc_*              parallel/series nesting, tanh/leaky/softmax convs, Lfys summarising, Addition,
                      1-augmented linear, strided/dilated convs, GroupNorm with ragged widths
  trained_gallicorpora  tests/resources/Gallicorpora+_best.safetensors (H=120, 3x13 / 3x9 kernels, 3 x BiLSTM-200, fp16 weights) on
                      the 29 bbox lines of input.webp through the reference's own legacy rpred (the path of
                      tests/test_tasks.py:117-130, criterion: SequenceMatcher ratio > 0.9 against box_rec.pkl); per line the
                      captured network input (stored as the uint8 pixels it was derived from + the inversion maximum, checked
                      to reproduce the float tensor bit for bit), logits, label tuples, prediction
  trained_arabic      tests/resources/all_arabic_scripts.safetensors on the bbox lines of arabic.webp (bboxes of
                      arabic_bbox_records.pkl), same content, right-to-left script
  trained_blla       kraken/blla.mlmodel (the shipped segmentation model, real weights) on one 3x1800x1350 page (BASELINE cfg3:
                      a 2400x3200 page after the reference's fixed resize): logits 4x450x338 fp32, sub-sampled sigmoid heat map.
                      The page is tests/golden/page_input.webp (a copy of the reference's test image input.webp) resized with PIL;
                      the fixture stores a checksum of the resulting tensor so that a different PIL build is noticed
  --trained           only regenerate the three cases above (python oracle/make_golden.py --trained)
Weights for the seeded cases come from OracleModel.init_like_reference(seed) (torch CPU generator,
deterministic for this torch build) and are NOT stored; the reference model is loaded with them.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import refshim  # noqa: E402

refshim.install()
warnings.simplefilter('ignore')

from collections import defaultdict  # noqa: E402

from PIL import Image  # noqa: E402

from kraken.containers import BBoxLine, Segmentation  # noqa: E402
from kraken.lib.ctc_decoder import greedy_decoder  # noqa: E402
from kraken.lib.models import TorchSeqRecognizer  # noqa: E402
from kraken.lib.vgsl.model import TorchVGSLModel  # noqa: E402
from kraken.rpred import mm_rpred, rpred  # noqa: E402

import vgsl_oracle as vo  # noqa: E402
from fixtures import blla_page_tensor, line_from_u8  # noqa: E402
from kraken_b200.weights import load_coreml  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
RES = os.path.join(refshim.REFERENCE_ROOT, 'tests', 'resources')

CFG2 = '[1,48,0,1 Cr3,3,32 Mp2,2 Cr3,3,64 Mp2,2 S1(1x0)1,3 Lbx256 O1c200]'
DEFAULT_REC = ('[1,120,0,1 Cr3,13,32 Do0.1,2 Mp2,2 Cr3,13,32 Do0.1,2 Mp2,2 Cr3,9,64 Do0.1,2 Mp2,2 Cr3,9,64 Do0.1,2 '
               'S1(1x0)1,3 Lbx200 Do0.1,2 Lbx200 Do0.1,2 Lbx200 Do O1c80]')
BLLA = ('[1,1800,0,3 Cr7,7,64,2,2 Gn32 Cr3,3,128,2,2 Gn32 Cr3,3,128 Gn32 Cr3,3,256 Gn32 Cr3,3,256 Gn32 '
        'Lbx32 Lby32 Cr1,1,32 Gn32 Lby32 Lbx32 O2l4]')
MISC = {
    'misc_parallel': '[1,48,0,1 Cr3,3,16 Mp2,2 ([Cr3,3,8 Ct1,1,8] I) S1(1x0)1,3 Lfx16 Lrx8 O1ca10]',
    'misc_summarize': '[1,32,0,1 Cr3,3,8 Mp2,2xyz A3,4 Lfys16 Lbx8 O1s7]',
    'misc_strided_gn': '[1,30,0,1 Cr3,3,32,2,2 Gn32 Cr3,3,64,2,2 Gn32 S1(1x0)1,3 O1c16]',
    'misc_acts': '[1,24,0,3 Clr5,3,8,1,2 Cm3,3,6 Cl3,5,4,1,1,2,2 Mp3,3,2,2 Cs1,1,5 Lby6 O2s3]',
    'misc_featseq': '[1,1,0,48 Lbx20 Do O1c59]',
}


def _dec_arrays(dec):
    """list[list[(label,start,end,conf)]] -> fixed-stride arrays."""
    n = len(dec)
    m = max([len(d) for d in dec] + [1])
    lab = np.zeros((n, m), np.int32)
    st = np.zeros((n, m), np.int32)
    en = np.zeros((n, m), np.int32)
    cf = np.zeros((n, m), np.float32)
    cnt = np.zeros(n, np.int32)
    for i, d in enumerate(dec):
        cnt[i] = len(d)
        for j, (l, s, e, c) in enumerate(d):
            lab[i, j], st[i, j], en[i, j], cf[i, j] = l, s, e, c
    return dict(dec_label=lab, dec_start=st, dec_end=en, dec_conf=cf, dec_count=cnt)


def _ref_model(spec, weights, codec=None):
    m = TorchVGSLModel(vgsl=spec, codec=codec) if codec else TorchVGSLModel(vgsl=spec)
    sd = {k: torch.as_tensor(np.asarray(v.detach() if torch.is_tensor(v) else v)).float() for k, v in weights.items()}
    m.load_state_dict(sd)
    m.eval()
    return m


def seeded_case(name, spec, seed, x, lens, seg_size=None, temperature=1.0):
    om = vo.OracleModel(spec)
    w = om.init_like_reference(seed)
    ref = _ref_model(spec, w)
    assert ref.user_metadata['vgsl'] == om.named_spec
    with torch.inference_mode():
        logits, olens = ref.nn(x, lens)
    d = dict(spec=spec, named_spec=om.named_spec, seed=seed, x=x.numpy(), logits=logits.numpy(),
             temperature=np.float32(temperature))
    if lens is not None:
        d['lens'] = lens.numpy().astype(np.int64)
        d['olens'] = olens.numpy().astype(np.int64)
    if logits.shape[2] == 1:
        probs = (logits / temperature).softmax(1).squeeze(2)
        ol = olens if olens is not None else torch.tensor([probs.shape[-1]] * probs.shape[0])
        d.update(_dec_arrays(greedy_decoder(probs, ol)))
        d['probs'] = probs.numpy()
    if seg_size is not None:
        hm = torch.sigmoid(torch.nn.functional.interpolate(logits, size=seg_size))
        d['heatmap'] = hm.numpy().astype(np.float16)       # halves the fixture; compared at 2e-3
        d['seg_size'] = np.asarray(seg_size, np.int64)
    # the oracle must agree bit-for-bit with the reference here (same machine, same ATen build)
    ol_logits, ol_olens = om.forward(x, lens)
    assert torch.equal(ol_logits, logits), name
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **d)
    print(f'{name}: logits {tuple(logits.shape)}  |max| {logits.abs().max():.3f}')


def overfit_cases():
    mf = load_coreml(os.path.join(RES, 'overfit.mlmodel'))[0]
    ref = _ref_model(mf.vgsl, mf.weights, codec=mf.codec)
    ref.user_metadata.update(mf.metadata)
    ref.one_channel_mode = mf.metadata.get('one_channel_mode')
    ref.seg_type = mf.metadata.get('seg_type')
    rec = TorchSeqRecognizer(ref, device='cpu')
    captured = {}
    orig_forward = rec.forward

    def spy(line, lens=None):
        captured['x'] = line.detach().clone()
        o = orig_forward(line, lens)
        with torch.inference_mode():
            captured['logits'] = rec.nn.nn(line, lens)[0].detach().clone()
        return o
    rec.forward = spy
    im = Image.open(os.path.join(RES, '000236.png'))
    seg = Segmentation(type='bbox', imagename='000236.png', lines=[BBoxLine(id='foo', bbox=[0, 0, 2544, 156])],
                       text_direction='horizontal-lr', script_detection=False)
    runs = {
        'cfg1_overfit_bbox': (lambda: rpred(rec, im, seg, True), 'ܡ ܘܡ ܗ ܡܕܐ ܐ ܐܐ ܡ ܗܗܐܐܐܕ'),
        'cfg1_overfit_nobidi': (lambda: mm_rpred(defaultdict(lambda: rec), im, seg, bidi_reordering=False),
                                'ܕܗܣܐܕ ܪܝ .ܡܡ ܐܠܠ ܗܠ ܐܘܗ ܟܘܗܢ ܡܡ ܐܠ'),
    }
    for name, (fn, expect) in runs.items():
        record = next(fn())
        assert record.prediction == expect, (name, record.prediction)
        x = captured['x']
        logits = captured['logits']
        probs = logits.softmax(1).squeeze(2)
        dec = greedy_decoder(probs, torch.tensor([probs.shape[-1]]))
        raw = ''.join(c for c, *_ in rec.codec.decode(dec[0]))
        d = dict(spec=mf.vgsl, x=x.numpy(), logits=logits.numpy(), probs=probs.numpy(),
                 prediction=expect, raw_prediction=raw, codec=json.dumps(mf.codec, ensure_ascii=False),
                 one_channel_mode=str(mf.metadata.get('one_channel_mode')), seg_type=str(mf.metadata.get('seg_type')),
                 cuts=np.asarray(record.cuts, np.int64), confidences=np.asarray(record.confidences, np.float32))
        d.update(_dec_arrays(dec))
        d.update({'w::' + k: v for k, v in mf.weights.items()})
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **d)
        print(f'{name}: input {tuple(x.shape)} -> "{record.prediction}"  raw "{raw}"')


def _fp16_exact(w):
    """fp16 storage for weights that are fp16 values anyway (halves the fixture); fp32 otherwise."""
    out = {}
    for k, v in w.items():
        v = np.asarray(v, np.float32)
        h = v.astype(np.float16)
        out['w::' + k] = h if np.array_equal(h.astype(np.float32), v) else v
    return out


def _u8_source(x):
    """The reference's transforms end in `ToDtype(float32, scale=True)` + `tensor_invert` (max - im): recover the uint8 pixels and
    the maximum so that the fixture stores a quarter of the bytes (and compresses well), and prove the round trip is exact."""
    for m255 in range(255, 0, -1):                            # the line's maximum pixel (255 whenever there is white padding)
        mx = torch.tensor(float(m255)).mul_(1.0 / 255)
        k = torch.round((mx - x) * 255.0)                     # x = max - pixel / 255
        if float(k.max()) > 255 or float(k.min()) < 0:
            continue
        u8src = k.to(torch.uint8)                             # original pixel values
        if torch.equal(line_from_u8(u8src[0].numpy()), x):    # ToDtype(scale) + tensor_invert, as the tests rebuild it
            break
    else:
        raise AssertionError('uint8 reconstruction of a captured line is not exact')
    return u8src.numpy()


def trained_case(name, model_file, image, seg, expected_by_id, pad):
    from difflib import SequenceMatcher
    from kraken_b200.weights import load_model_file
    mf = load_model_file(os.path.join(RES, model_file))[0]
    ref = _ref_model(mf.vgsl, mf.weights, codec=mf.codec)
    ref.user_metadata.update(mf.metadata)
    ref.one_channel_mode = mf.metadata.get('one_channel_mode')
    ref.seg_type = mf.metadata.get('seg_type')
    rec = TorchSeqRecognizer(ref, device='cpu')
    cap = []
    orig_forward = rec.forward

    def spy(line, lens=None):
        o = orig_forward(line, lens)
        with torch.inference_mode():
            cap.append((line.detach().clone(), rec.nn.nn(line, lens)[0].detach().clone()))
        return o
    rec.forward = spy
    im = Image.open(os.path.join(RES, image))
    records = list(rpred(rec, im, seg, pad=pad))
    assert len(records) == len(cap) == len(seg.lines), (len(records), len(cap))
    d = dict(spec=mf.vgsl, codec=json.dumps(mf.codec, ensure_ascii=False), n_lines=np.int64(len(cap)), pad=np.int64(pad),
             one_channel_mode=str(mf.metadata.get('one_channel_mode')), seg_type=str(mf.metadata.get('seg_type')))
    d.update(_fp16_exact(mf.weights))
    ok = 0
    for i, ((x, logits), r) in enumerate(zip(cap, records)):
        probs = logits.softmax(1).squeeze(2)
        dec = greedy_decoder(probs, torch.tensor([probs.shape[-1]]))
        raw = ''.join(c for c, *_ in rec.codec.decode(dec[0]))
        d[f'u8::{i}'] = _u8_source(x)[0]                 # (1, H, W) uint8
        if i < 8:
            d[f'logits::{i}'] = logits.numpy()
        a = _dec_arrays(dec)
        for k, v in a.items():
            d[f'{k}::{i}'] = v[0]
        d[f'raw::{i}'] = raw
        d[f'pred::{i}'] = r.prediction
        if expected_by_id is not None:
            exp = expected_by_id[i]
            d[f'expected::{i}'] = exp
            ok += SequenceMatcher(isjunk=None, a=r.prediction, b=exp).ratio() > 0.9
    if expected_by_id is not None:
        assert ok == len(cap), f'{name}: only {ok}/{len(cap)} lines meet the reference test criterion'
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **d)
    ws = sorted(int(x.shape[-1]) for x, _ in cap)
    print(f'{name}: {len(cap)} lines, widths {ws[0]}..{ws[-1]}, criterion lines ok: {ok}')


def trained_cases():
    import pickle
    with open(os.path.join(RES, 'box_rec.pkl'), 'rb') as fp:
        box = pickle.load(fp)
    seg = Segmentation(type='bbox', imagename='input.webp', lines=[BBoxLine(id=l.id, bbox=l.bbox) for l in box.lines],
                       text_direction='horizontal-lr', script_detection=False)
    trained_case('trained_gallicorpora', 'Gallicorpora+_best.safetensors', 'input.webp', seg, [l.prediction for l in box.lines], pad=16)
    with open(os.path.join(RES, 'arabic_bbox_records.pkl'), 'rb') as fp:
        ar = pickle.load(fp)
    seg = Segmentation(type='bbox', imagename='arabic.webp', lines=[BBoxLine(id=f'l{i}', bbox=list(r.bbox)) for i, r in enumerate(ar)],
                       text_direction='horizontal-lr', script_detection=False)
    trained_case('trained_arabic', 'all_arabic_scripts.safetensors', 'arabic.webp', seg, None, pad=16)


def blla_real_case():
    import shutil
    import zlib
    from kraken_b200.weights import load_model_file
    mf = load_model_file(os.path.join(refshim.REFERENCE_ROOT, 'kraken', 'blla.mlmodel'))[0]
    ref = _ref_model(mf.vgsl, mf.weights)
    page_copy = os.path.join(OUT, 'page_input.webp')
    if not os.path.exists(page_copy):
        shutil.copyfile(os.path.join(RES, 'input.webp'), page_copy)
    x = blla_page_tensor(page_copy)
    with torch.inference_mode():
        logits, _ = ref.nn(x, None)
        hm = torch.sigmoid(torch.nn.functional.interpolate(logits, size=(1800, 1350)))
    om = vo.OracleModel(mf.vgsl, {k: torch.as_tensor(v) for k, v in mf.weights.items()})
    ol, _ = om.forward(x, None)
    assert torch.equal(ol, logits)
    d = dict(spec=mf.vgsl, logits=logits.numpy(), heatmap_sub=hm[:, :, 3::7, 2::7].numpy().astype(np.float16),
             x_crc=np.int64(zlib.crc32(x.numpy().tobytes())), class_mapping=json.dumps(mf.metadata.get('class_mapping')))
    d.update({'w::' + k: np.asarray(v, np.float32) for k, v in mf.weights.items()})
    np.savez_compressed(os.path.join(OUT, 'trained_blla.npz'), **d)
    print(f'trained_blla: logits {tuple(logits.shape)} |max| {logits.abs().max():.3f}, heat map range {float(hm.min()):.3g}..{float(hm.max()):.3g}')


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if '--trained' in sys.argv:
        trained_cases()
        blla_real_case()
        return
    overfit_cases()
    g = torch.Generator().manual_seed(1234)
    # ragged batch incl. degenerate widths: 1 column, 3 columns (one output step), odd widths
    lens = torch.tensor([160, 101, 37, 5, 158], dtype=torch.long)
    x = torch.rand(5, 1, 48, 160, generator=g)
    for i, l in enumerate(lens.tolist()):
        x[i, ..., l:] = 0
    seeded_case('cfg2_small', CFG2, 0, x, lens)
    seeded_case('cfg2_small_nolens', CFG2, 1, torch.rand(2, 1, 48, 96, generator=g), None, temperature=2.0)
    lens = torch.tensor([200, 133, 64], dtype=torch.long)
    x = torch.rand(3, 1, 120, 200, generator=g)
    for i, l in enumerate(lens.tolist()):
        x[i, ..., l:] = 0
    seeded_case('rec_default_small', DEFAULT_REC, 2, x, lens)
    seeded_case('seg_blla_small', BLLA, 3, torch.rand(2, 3, 96, 72, generator=g), None, seg_size=(90, 70))
    lens3 = torch.tensor([96, 50, 17], dtype=torch.long)
    for name, spec in MISC.items():
        om = vo.OracleModel(spec)
        b, c, h, w = om.input
        x = torch.rand(3, c, h if h else 1, 96, generator=g)
        try:
            seeded_case(name + '_lens', spec, 5, x, lens3)
        except Exception as e:                          # e.g. Lby under seq_lens -> reference raises
            print(f'{name}_lens: reference raises {type(e).__name__}: {e}')
        seeded_case(name, spec, 5, x, None)
    trained_cases()
    blla_real_case()
    sizes = {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))}
    print(json.dumps(sizes, indent=1), sum(sizes.values()))


if __name__ == '__main__':
    main()
