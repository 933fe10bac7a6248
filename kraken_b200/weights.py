"""
Native readers for kraken's two on-disk model formats - no coremltools, no safetensors package.

Replaces, for the engine, the weight-extraction half of
  kraken/models/loaders.py:46-151   (load_safetensors: `kraken_meta` JSON in the header, `<uuid>.` key prefix)
  kraken/models/loaders.py:153-254  (load_coreml: `vgsl` / `codec` / `kraken_meta` user metadata)
  kraken/models/_coreml.py:10-108   (conv / LSTM / innerProduct / groupnorm tensor views)

Both return `ModelFile(vgsl, codec, metadata, weights)` where `weights` maps the reference's state-dict
keys (`nn.<name>.{co,lin,layer}.*`) to fp32 numpy arrays (fp16 storage is widened, as the reference's
loader ends up with fp32 parameters: tests/test_loaders.py:117-149).
"""
from __future__ import annotations

import ast
import json
import struct
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

__all__ = ['ModelFile', 'load_safetensors', 'load_coreml', 'load_model_file']


@dataclass
class ModelFile:
    vgsl: str
    codec: Optional[dict]
    metadata: dict = field(default_factory=dict)
    weights: dict = field(default_factory=dict)


# --------------------------------------------------------------------------------------
# safetensors
# --------------------------------------------------------------------------------------
_ST_DTYPES = {'F32': np.float32, 'F16': np.float16, 'F64': np.float64, 'I64': np.int64, 'I32': np.int32,
              'I16': np.int16, 'I8': np.int8, 'U8': np.uint8, 'BOOL': np.bool_}


def _bf16_to_f32(raw: bytes, shape):
    u = np.frombuffer(raw, dtype=np.uint16).astype(np.uint32) << 16
    return u.view(np.float32).reshape(shape)


def _maybe_json(v):
    if isinstance(v, str):
        try:
            return json.loads(v)
        except Exception:
            try:
                return ast.literal_eval(v)
            except Exception:
                return v
    return v


def load_safetensors(path: str, tasks=None) -> list[ModelFile]:
    with open(path, 'rb') as fh:
        head = fh.read(8)
        if len(head) != 8:
            raise ValueError(f'Invalid safetensors file {path}: truncated header')
        (n,) = struct.unpack('<Q', head)
        if n > (1 << 30):
            raise ValueError(f'Invalid safetensors file {path}: implausible header size {n}')
        try:
            header = json.loads(fh.read(n))
        except Exception as e:
            raise ValueError(f'Invalid safetensors file {path}: {e}') from e
        blob = fh.read()
    meta = header.pop('__metadata__', None)
    if meta is None:
        raise ValueError(f'No model metadata found in {path}.')
    try:
        model_map = json.loads(meta.get('kraken_meta', 'null'))
    except json.JSONDecodeError as e:
        raise ValueError(f'Invalid `kraken_meta` JSON in {path}: {e}') from e
    if not isinstance(model_map, dict):
        raise ValueError(f'Invalid `kraken_meta` metadata in {path}: expected object, got {type(model_map).__name__}.')
    out = []
    for prefix, data in model_map.items():
        if not isinstance(data, dict):
            raise ValueError(f'Invalid metadata for model `{prefix}` in {path}: expected object, got {type(data).__name__}.')
        mtasks = data.get('_tasks') or []
        if tasks and not set(tasks).intersection(mtasks):
            continue
        if data.get('_model') != 'TorchVGSLModel':
            continue                                   # other model classes are not the engine's business
        md = {k: v for k, v in data.items() if k not in ('_tasks', '_kraken_min_version', '_model')}
        md['model_type'] = list(mtasks)
        vgsl = md.pop('vgsl', None)
        if not vgsl:
            raise ValueError(f'No VGSL spec in model metadata for {path}')
        codec = _maybe_json(md.pop('codec', None))
        for k in ('hyper_params', 'accuracy', 'class_mapping'):
            if k in md:
                md[k] = _maybe_json(md[k])
        weights = {}
        pre = prefix + '.'
        for key, info in header.items():
            if not key.startswith(pre):
                continue
            lo, hi = info['data_offsets']
            raw = blob[lo:hi]
            if info['dtype'] == 'BF16':
                arr = _bf16_to_f32(raw, info['shape'])
            else:
                arr = np.frombuffer(raw, dtype=_ST_DTYPES[info['dtype']]).reshape(info['shape'])
            weights[key[len(pre):]] = np.ascontiguousarray(arr, dtype=np.float32) if arr.dtype.kind == 'f' else arr.copy()
        out.append(ModelFile(vgsl, codec, md, weights))
    return out


# --------------------------------------------------------------------------------------
# CoreML .mlmodel: wire-level protobuf walk (only the fields kraken writes)
# --------------------------------------------------------------------------------------
def _varint(buf: memoryview, i: int):
    r = 0
    s = 0
    while True:
        b = buf[i]
        i += 1
        r |= (b & 0x7F) << s
        if not b & 0x80:
            return r, i
        s += 7


def _fields(buf):
    """Yields (field_number, wire_type, value) - value is int for varint/fixed, memoryview for bytes."""
    buf = memoryview(buf)
    i, n = 0, len(buf)
    while i < n:
        key, i = _varint(buf, i)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(buf, i)
        elif wt == 1:
            v = bytes(buf[i:i + 8])
            i += 8
        elif wt == 2:
            ln, i = _varint(buf, i)
            v = buf[i:i + ln]
            i += ln
        elif wt == 5:
            v = bytes(buf[i:i + 4])
            i += 4
        else:
            raise ValueError(f'Failure parsing model protobuf: unsupported wire type {wt}')
        yield fno, wt, v


def _first(buf, fno):
    for f, _, v in _fields(buf):
        if f == fno:
            return v
    return None


def _all(buf, fno):
    return [v for f, _, v in _fields(buf) if f == fno]


def _floats(weight_params) -> np.ndarray:
    """WeightParams{1: repeated float floatValue (packed), 2: bytes float16Value}."""
    if weight_params is None:
        return np.zeros(0, np.float32)
    chunks = []
    for f, wt, v in _fields(weight_params):
        if f == 1 and wt == 2:
            chunks.append(np.frombuffer(bytes(v), dtype='<f4'))
        elif f == 1 and wt == 5:
            chunks.append(np.frombuffer(v, dtype='<f4'))
        elif f == 2 and wt == 2:
            chunks.append(np.frombuffer(bytes(v), dtype='<f2').astype(np.float32))
    return np.concatenate(chunks).astype(np.float32) if chunks else np.zeros(0, np.float32)


def _uints(buf, fno):
    out = []
    for f, wt, v in _fields(buf):
        if f == fno:
            if wt == 0:
                out.append(v)
            else:
                vv = memoryview(v)
                i = 0
                while i < len(vv):
                    x, i = _varint(vv, i)
                    out.append(x)
    return out


def _str(v) -> str:
    return bytes(v).decode('utf-8') if v is not None else ''


def _lstm_dir(params, input_size, hidden, name, sfx, weights):
    wx = [_floats(_first(params, i)) for i in (1, 2, 3, 4)]       # W_x: i, f, z(g), o
    wh = [_floats(_first(params, i)) for i in (20, 21, 22, 23)]   # W_h
    bs = [_floats(_first(params, i)) for i in (40, 41, 42, 43)]
    weights[f'nn.{name}.layer.weight_ih_l0{sfx}'] = np.stack(wx).reshape(-1, input_size)
    weights[f'nn.{name}.layer.weight_hh_l0{sfx}'] = np.stack(wh).reshape(-1, hidden)
    b = np.stack(bs).reshape(-1) if all(len(x) for x in bs) else np.zeros(4 * hidden, np.float32)
    weights[f'nn.{name}.layer.bias_hh_l0{sfx}'] = b
    weights[f'nn.{name}.layer.bias_ih_l0{sfx}'] = np.zeros_like(b)       # CoreML stores one bias (_coreml.py:50-52)


def _strip_suffix(s, suf):
    return s[:-len(suf)] if s.endswith(suf) else s


def load_coreml(path: str, tasks=None) -> list[ModelFile]:
    with open(path, 'rb') as fh:
        blob = fh.read()
    try:
        desc = _first(blob, 2)
        nnet = _first(blob, 500)
        if desc is None or nnet is None:
            raise ValueError('not a NeuralNetwork CoreML model')
        user = {}
        md_msg = _first(desc, 100)
        if md_msg is not None:
            for entry in _all(md_msg, 100):
                user[_str(_first(entry, 1))] = _str(_first(entry, 2))
    except (IndexError, ValueError) as e:
        raise ValueError(f'Failure parsing model protobuf: {e}') from e

    has_meta = 'kraken_meta' in user
    try:
        metadata = json.loads(user.get('kraken_meta', '{}'))
    except json.JSONDecodeError as e:
        raise ValueError(f'Invalid `kraken_meta` JSON in {path}: {e}') from e
    if not isinstance(metadata, dict):
        raise ValueError(f'Invalid `kraken_meta` metadata in {path}: expected object, got {type(metadata).__name__}.')
    mt = metadata.get('model_type')
    if isinstance(mt, str):
        mt = [mt] if mt else []
    if not isinstance(mt, list) or not mt or not all(isinstance(x, str) and x for x in mt):
        if has_meta:
            raise ValueError(f'Invalid `model_type` metadata in {path}: expected string or list[str], got {type(mt).__name__}.')
        mt = ['recognition']                       # pre-kraken_meta files are recognisers (loaders.py:199-203)
    metadata['model_type'] = mt
    vgsl = user.get('vgsl') or metadata.get('vgsl')
    metadata.pop('codec', None)
    metadata.pop('vgsl', None)
    if not vgsl:
        raise ValueError(f'No VGSL spec in model metadata for {path}')
    if tasks and not set(mt).intersection(tasks):
        return []
    codec = json.loads(user.get('codec', 'null'))

    weights: dict = {}
    for layer in _all(nnet, 1):
        name = _str(_first(layer, 1))
        for fno, wt, body in _fields(layer):
            if wt != 2:
                continue
            if fno == 100:                                           # ConvolutionLayerParams
                nm = _strip_suffix(name, '_conv')
                out_c, k_c = _first(body, 1), _first(body, 2)
                ks = _uints(body, 20)
                deconv = bool(_first(body, 60) or 0)
                w = _floats(_first(body, 90))
                shape = (k_c, out_c, *ks) if deconv else (out_c, k_c, *ks)
                weights[f'nn.{nm}.co.weight'] = w.reshape(shape)
                weights[f'nn.{nm}.co.bias'] = _floats(_first(body, 91))
            elif fno == 140:                                         # InnerProductLayerParams
                nm = _strip_suffix(name, '_lin')
                in_c, out_c = _first(body, 1), _first(body, 2)
                weights[f'nn.{nm}.lin.weight'] = _floats(_first(body, 20)).reshape(out_c, in_c)
                weights[f'nn.{nm}.lin.bias'] = _floats(_first(body, 21))
            elif fno in (420, 430):                                  # uni / bi-directional LSTM
                nm = _strip_suffix(name, '_transposed')
                in_sz, hid = _first(body, 1), _first(body, 2)
                wps = _all(body, 20)
                _lstm_dir(wps[0], in_sz, hid, nm, '', weights)
                if fno == 430 and len(wps) > 1:
                    _lstm_dir(wps[1], in_sz, hid, nm, '_reverse', weights)
            elif fno == 500:                                         # CustomLayerParams
                if _str(_first(body, 10)) == 'groupnorm':
                    ws = _all(body, 20)
                    weights[f'nn.{name}.layer.weight'] = _floats(ws[0])
                    weights[f'nn.{name}.layer.bias'] = _floats(ws[1])
    for k in ('hyper_params', 'accuracy', 'class_mapping'):
        if k in metadata:
            metadata[k] = _maybe_json(metadata[k])
    return [ModelFile(vgsl, codec, metadata, {k: np.ascontiguousarray(v, np.float32) for k, v in weights.items()})]


def load_model_file(path: str, tasks=None) -> list[ModelFile]:
    """Tries the formats in turn; ValueError means 'not mine' (the convention of loaders.py:35-43)."""
    errs = []
    for fn in (load_safetensors, load_coreml):
        try:
            models = fn(path, tasks)
            return models
        except (ValueError, KeyError, struct.error, UnicodeDecodeError) as e:
            errs.append(f'{fn.__name__}: {e}')
    raise ValueError(f'No loader found for {path}: ' + '; '.join(errs))
