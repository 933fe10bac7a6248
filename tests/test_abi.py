"""The C-ABI library loads on a CPU-only box and exports every symbol include/kraken_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'kraken_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(kb_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ('kb_model_create', 'kb_model_load_tensor', 'kb_model_finalize', 'kb_forward', 'kb_recognize',
                 'kb_ctc_greedy_decode', 'kb_segment', 'kb_last_error'):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import kraken_b200
    lib = ctypes.CDLL(kraken_b200.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    # and the ctypes table covers the header, so nothing can be called with default (int) signatures
    from kraken_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()


def test_library_was_built_from_the_sources_in_the_tree():
    """The .so is git-ignored but travels to the GPU box with the snapshot: a stale binary must not pass for the current kernels."""
    import __graft_entry__ as ge
    import kraken_b200
    assert kraken_b200.lib.kb_source_hash().decode() == ge.source_hash()


def test_abi_version_and_device_count():
    import kraken_b200
    assert kraken_b200.lib.kb_abi_version() == 4
    assert kraken_b200.device_count() >= 0


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must fail loudly (EngineError), never compute on the CPU."""
    import torch
    import kraken_b200 as kb
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    m = kb.TorchVGSLModel(vgsl='[1,48,0,1 Cr3,3,8 Mp2,2 S1(1x0)1,3 Lbx8 O1c5]')
    with pytest.raises(kb.EngineError):
        m.nn(torch.rand(1, 1, 48, 32))
    with pytest.raises(kb.EngineError):
        kb.greedy_decoder(torch.rand(5, 10))
    with pytest.raises(kb.EngineError):
        kb.TorchSeqRecognizer(m, device='cuda:0')


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'kraken_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.hpp', '.h')):
                txt = open(os.path.join(dirpath, f), errors='ignore').read()
                assert 'vgsl_oracle' not in txt and 'refshim' not in txt and 'np_kernels' not in txt, f

