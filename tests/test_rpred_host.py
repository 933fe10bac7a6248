"""Host-side batching / routing helpers of kraken_b200.rpred (no GPU): the padded-batch semantics parity is judged on
(kraken/lib/vgsl/rpred.py:129-131) and mm_rpred's tag resolution (kraken/rpred.py:373-391)."""
from collections import defaultdict

import pytest
import torch

from kraken_b200.rpred import pad_batch, resolve_type_to_model


def test_pad_batch_right_pads_with_zeros_and_keeps_widths():
    lines = [torch.rand(1, 48, w) + 0.5 for w in (37, 120, 5)]
    seqs, lens = pad_batch(lines)
    assert tuple(seqs.shape) == (3, 1, 48, 120) and lens.dtype == torch.int64 and lens.tolist() == [37, 120, 5]
    for i, l in enumerate(lines):
        w = l.shape[2]
        assert torch.equal(seqs[i, :, :, :w], l)
        assert float(seqs[i, :, :, w:].abs().sum()) == 0.0
    # identical to the reference's expression
    ref = torch.stack([torch.nn.functional.pad(l, pad=(0, 120 - l.shape[2])) for l in lines])
    assert torch.equal(seqs, ref)


def test_resolve_type_to_model():
    nets = {'latin': 'A', 'greek': 'B'}
    assert resolve_type_to_model('greek', nets) == ('greek', 'B')
    assert resolve_type_to_model(None, nets, default='D') == ('default', 'D')
    assert resolve_type_to_model('syriac', nets, default='D') == ('syriac', 'D')
    with pytest.raises(KeyError):
        resolve_type_to_model('syriac', nets)
    dd = defaultdict(lambda: 'D', nets)
    assert resolve_type_to_model('latin', dd, dd.default_factory()) == ('latin', 'A')


def test_legacy_segmentation_mask_is_the_references_boolean_indexing():
    """kraken/blla.py:115 `tensor_im[~transforms(mask).bool()] = 0` on a one-channel page; shape errors as the reference's indexing raises"""
    import pytest
    import torch
    from kraken_b200.blla import apply_legacy_mask
    g = torch.Generator().manual_seed(0)
    page = torch.rand(1, 40, 30, generator=g)
    mask = (torch.rand(1, 40, 30, generator=g) > 0.5).float() * 0.7          # any non-zero value keeps a pixel
    ref = page.clone()
    ref[~mask.bool()] = 0
    assert torch.equal(apply_legacy_mask(page, mask), ref)
    with pytest.raises(IndexError):
        apply_legacy_mask(torch.rand(3, 40, 30), mask)
