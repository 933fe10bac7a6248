"""kraken_b200.codec.PytorchCodec (mirror of kraken/lib/codec.py) - pure semantics, and identity with the reference's class on random
label streams when the reference checkout is mounted (build container only)."""
import os
import random
import sys

import pytest

from kraken_b200.codec import KrakenCodecException, KrakenEncodeException, PytorchCodec

HAVE_REF = os.path.isdir('/root/reference/kraken')


def test_charset_forms_and_validity():
    c = PytorchCodec('cba')                                   # string: sorted, labels from 1
    assert c.c2l == {'a': [1], 'b': [2], 'c': [3]} and len(c) == 3 and c.max_label == 3
    assert PytorchCodec(['ab', 'c']).c2l == {'ab': [1], 'c': [2]}
    with pytest.raises(KrakenCodecException):
        PytorchCodec('aab')                                   # duplicate entry
    with pytest.raises(KrakenCodecException):
        PytorchCodec({'a': [1], 'b': [1]})                    # two code points, one label sequence
    with pytest.raises(KrakenCodecException):
        PytorchCodec({'a': [1], 'b': [1, 2]})                 # not prefix free


def test_encode_prefers_longest_match_and_skips_unknown():
    c = PytorchCodec({'a': [1], 'ab': [2, 3], 'b': [4]})
    assert c.encode('abab').tolist() == [2, 3, 2, 3]
    assert c.encode('ba?a').tolist() == [4, 1, 1]             # unknown code point dropped
    with pytest.raises(KrakenEncodeException):
        PytorchCodec({'a': [1]}, strict=True).encode('ab')


def test_decode_single_and_multi_label_codes():
    c = PytorchCodec({'a': [1], 'xy': [2, 3], 'b': [4]})
    dec = c.decode([(1, 0, 1, 0.5), (2, 2, 3, 0.2), (3, 4, 6, 0.6), (4, 7, 8, 1.0), (9, 9, 9, 0.1)])
    assert [d[0] for d in dec] == ['a', 'x', 'y', 'b']        # undecodable label 9 dropped (non-strict)
    assert dec[1][1:3] == (2, 6) and dec[2][1:3] == (2, 6)    # multi-label code: first start, last end
    assert abs(dec[1][3] - 0.4) < 1e-12                       # ... mean confidence
    with pytest.raises(KrakenEncodeException):
        PytorchCodec({'a': [1]}, strict=True).decode([(7, 0, 0, 1.0)])


@pytest.mark.skipif(not HAVE_REF, reason='reference checkout not mounted')
def test_identical_to_reference_codec_on_random_streams():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
    import refshim
    refshim.install()
    from kraken.lib.codec import PytorchCodec as RefCodec
    rnd = random.Random(5)
    charsets = [{'a': [1], 'b': [2], 'c': [3]},
                {'a': [1], 'ab': [2, 3], 'b': [4], 'cde': [5, 6, 7]},
                {chr(0x710 + i): [i + 1] for i in range(15)}]                 # cfg1's Syriac alphabet shape
    for cs in charsets:
        ours, ref = PytorchCodec(cs), RefCodec(cs)
        assert len(ours) == len(ref) and ours.max_label == ref.max_label
        alphabet = ''.join(cs.keys()) + '?'
        for _ in range(50):
            s = ''.join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 12)))
            assert ours.encode(s).tolist() == ref.encode(s).tolist(), s
            labs = [(rnd.randint(1, ours.max_label + 1), 3 * i, 3 * i + rnd.randint(0, 2), rnd.random()) for i in range(rnd.randint(0, 10))]
            a, b = ours.decode(labs), ref.decode(labs)
            assert [x[:3] for x in a] == [tuple(y[:3]) for y in b]
            assert all(abs(x[3] - float(y[3])) < 1e-9 for x, y in zip(a, b))


@pytest.mark.parametrize('charset', [{'a': [1], 'b': [2], 'c': [4]},                      # 1:1 -> table lookup
                                     {'a': [1], 'xy': [2, 3], 'b': [4]},                  # multi-label code -> generic path
                                     {chr(0x710 + i): [i + 1] for i in range(15)}])
def test_decode_blocks_equals_decode(charset):
    import numpy as np
    rnd = np.random.default_rng(3)
    c = PytorchCodec(charset)
    n, t = 9, 40
    labels = rnd.integers(0, c.max_label + 3, (n, t)).astype(np.int32)       # includes undecodable labels
    starts = np.cumsum(rnd.integers(1, 4, (n, t)), axis=1).astype(np.int32)
    ends = (starts + rnd.integers(0, 3, (n, t))).astype(np.int32)
    confs = rnd.random((n, t)).astype(np.float32)
    counts = rnd.integers(0, t + 1, n).astype(np.int32)
    counts[0] = 0
    got = c.decode_blocks(labels, starts, ends, confs, counts)
    for i in range(n):
        ref = c.decode([(int(labels[i, j]), int(starts[i, j]), int(ends[i, j]), float(confs[i, j])) for j in range(counts[i])])
        text, s_, e_, cf = got[i]
        assert text == ''.join(r[0] for r in ref)
        assert s_.tolist() == [r[1] for r in ref] and e_.tolist() == [r[2] for r in ref]
        assert np.allclose(cf, [r[3] for r in ref], atol=1e-7)


def test_decode_blocks_strict():
    import numpy as np
    c = PytorchCodec({'a': [1], 'b': [2]}, strict=True)
    ok = c.decode_blocks(np.array([[1, 2, 1]]), np.zeros((1, 3), np.int32), np.zeros((1, 3), np.int32), np.ones((1, 3), np.float32), np.array([3]))
    assert ok[0][0] == 'aba'
    with pytest.raises(KrakenEncodeException):
        c.decode_blocks(np.array([[1, 7, 1]]), np.zeros((1, 3), np.int32), np.zeros((1, 3), np.int32), np.ones((1, 3), np.float32), np.array([3]))
