"""
ctypes binding of libkraken_b200.so (include/kraken_b200.h).

There is no CPU fallback anywhere in this package: if the shared library has not been built
(`python -c "import __graft_entry__ as g; g.build()"`) importing this module raises, and every
compute entry point raises `EngineError` when no Blackwell GPU is usable.
"""
from __future__ import annotations

import ctypes as C
import os

__all__ = ['lib', 'EngineError', 'KrakenInputException', 'check', 'LIB_PATH', 'LayerInfo']

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libkraken_b200.so')

KB_OK, KB_ERR_SPEC, KB_ERR_ARG, KB_ERR_CUDA, KB_ERR_UNSUPPORTED, KB_ERR_STATE, KB_ERR_SHAPE = range(7)


class EngineError(RuntimeError):
    """CUDA / engine-state failure (no GPU, launch failure, ...)."""


class KrakenInputException(Exception):
    """Mirror of kraken.lib.exceptions.KrakenInputException (raised for non-1 output height, models.py:113-114)."""


class LayerInfo(C.Structure):
    _fields_ = [('kind', C.c_int32), ('out_shape', C.c_int32 * 4), ('name', C.c_char * 64),
                ('path', C.c_char * 256), ('block', C.c_char * 128)]


if not os.path.exists(LIB_PATH):
    raise ImportError(f'{LIB_PATH} is missing - build it with `python -c "import __graft_entry__ as g; g.build()"` '
                      '(nvcc, sm_100a). kraken_b200 has no pure-Python or CPU fallback.')

lib = C.CDLL(LIB_PATH)

_vp, _i32, _i64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_pi32, _pf = C.POINTER(C.c_int32), C.POINTER(C.c_float)

_SIGS = {
    'kb_abi_version': (C.c_int, []),
    'kb_last_error': (C.c_char_p, []),
    'kb_source_hash': (C.c_char_p, []),
    'kb_device_count': (C.c_int, []),
    'kb_model_create': (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    'kb_model_destroy': (None, [_vp]),
    'kb_model_named_spec': (C.c_int, [_vp, C.c_char_p, C.c_size_t]),
    'kb_model_input_shape': (C.c_int, [_vp, _pi32]),
    'kb_model_output_shape': (C.c_int, [_vp, _pi32]),
    'kb_model_num_layers': (C.c_int, [_vp]),
    'kb_model_layer_info': (C.c_int, [_vp, C.c_int, C.POINTER(LayerInfo)]),
    'kb_model_num_tensors': (C.c_int, [_vp]),
    'kb_model_tensor_info': (C.c_int, [_vp, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(_i64), _pi32]),
    'kb_model_infer_dims': (C.c_int, [_vp, _i32, _i32, _i32, _pi32]),
    'kb_model_infer_lens': (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    'kb_model_load_tensor': (C.c_int, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i32]),
    'kb_model_finalize': (C.c_int, [_vp, C.c_int]),
    'kb_model_device': (C.c_int, [_vp]),
    'kb_forward': (C.c_int, [_vp, _vp, C.c_int, _i32, _i32, _i32, _vp, _vp, C.c_int, _vp, _vp]),
    'kb_recognize': (C.c_int, [_vp, _vp, C.c_int, _i32, _i32, _i32, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, C.c_int, _vp]),
    'kb_recognize_u8': (C.c_int, [_vp, _vp, C.c_int, _i32, _i32, _i32, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, C.c_int, _vp]),
    'kb_model_set_codec': (C.c_int, [_vp, _vp, _i32]),
    'kb_recognize_records': (C.c_int, [_vp, _vp, C.c_int, C.c_int, _i32, _i32, _i32, _vp, _vp, _f, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    'kb_forced_align': (C.c_int, [_vp, _vp, C.c_int, C.c_int, _i32, _i32, _i32, _vp, _vp, _f, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    'kb_forced_align_probs': (C.c_int, [_vp, C.c_int, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, C.c_int, _vp]),
    'kb_line_width': (_i32, [_i32, _i32, _i32, _i32]),
    'kb_prepare_lines_u8': (C.c_int, [_vp, _vp, C.c_int, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp, _vp, _vp]),
    'kb_set_pipeline_depth': (C.c_int, [_vp, _i32]),
    'kb_pipeline_depth': (C.c_int, [_vp]),
    'kb_recognize_async': (C.c_int, [_vp, _vp, C.c_int, C.c_int, _i32, _i32, _i32, _vp, _vp, _f, _i32, _vp, C.POINTER(_i64)]),
    'kb_wait': (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    'kb_ctc_greedy_decode': (C.c_int, [_vp, C.c_int, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, C.c_int, _vp]),
    'kb_segment': (C.c_int, [_vp, _vp, C.c_int, _i32, _i32, _i32, _i32, _i32, _vp, C.c_int, _vp]),
    'kb_debug_layer_output': (C.c_int, [_vp, C.c_char_p, _pi32, _vp, C.c_int]),
    'kb_debug_gemm': (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, C.c_int, C.c_int]),
    'kb_debug_axis_coeffs': (C.c_int, [_i32, _i32, _pi32, _vp, _vp, _i32]),
    'kb_launch_count': (_i64, [_vp]),
    'kb_range_fallback_count': (_i64, [_vp]),
    'kb_reset_launch_count': (None, [_vp]),
    'kb_set_timing': (C.c_int, [_vp, C.c_int]),
    'kb_timing_count': (C.c_int, [_vp]),
    'kb_timing_entry': (C.c_int, [_vp, C.c_int, C.c_char_p, C.c_size_t, _pf]),
}
EXPORTS = tuple(_SIGS)
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)          # AttributeError here == header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return (lib.kb_last_error() or b'').decode('utf-8', 'replace')


def check(rc: int):
    """Maps engine status codes onto the reference's exception types (include/kraken_b200.h)."""
    if rc == KB_OK:
        return
    msg = last_error()
    if rc == KB_ERR_SPEC:
        raise ValueError(msg)
    if rc == KB_ERR_SHAPE:
        raise KrakenInputException(msg)
    if rc == KB_ERR_ARG:
        raise ValueError(msg)
    if rc == KB_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise EngineError(msg)
