"""
kraken_b200 - Blackwell-native (sm_100a) engine for kraken's rpred / blla-forward hot path.

Importing the package loads `libkraken_b200.so` (C ABI: include/kraken_b200.h) and fails loudly if it has not been
built; there is no CPU or pure-PyTorch fallback.
"""
from ._lib import EngineError, KrakenInputException, LIB_PATH, lib  # noqa: F401
from .codec import PytorchCodec  # noqa: F401
from .ctc_decoder import greedy_decoder  # noqa: F401
from .models import TorchSeqRecognizer, load_any  # noqa: F401
from .vgsl import TorchVGSLModel  # noqa: F401

__version__ = '0.1.0'


def device_count() -> int:
    return int(lib.kb_device_count())
