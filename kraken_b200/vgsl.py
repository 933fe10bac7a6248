"""
Host-side mirror of kraken's `TorchVGSLModel` for the engine (reference: kraken/lib/vgsl/model.py:78-902).

Same names, argument meaning and error behaviour as the reference for the inference surface:
`input`, `output`, `spec`, `named_spec`, `user_metadata`, `codec`, `nn(x, seq_lens)`, `forward`,
`state_dict`/`load_state_dict` with the reference's keys, `init_weights`, `add_codec`, `resize_output`,
`append`, `load_model`, `to`, `eval`, the metadata properties, `prepare_for_inference` and `predict`.
The graph itself lives behind the C ABI (include/kraken_b200.h): the VGSL spec is parsed by the library,
weights are handed over with `kb_model_load_tensor`, and `nn` is one `kb_forward` call.  Training-only
members (`criterion`, autograd) are intentionally absent - no backward pass is in scope.
"""
from __future__ import annotations

import ctypes as C
import json
import re
from typing import Any, Iterable, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .codec import PytorchCodec

__all__ = ['TorchVGSLModel', 'EngineNet']

_KIND = {1: 'conv', 2: 'maxpool', 3: 'reshape', 4: 'rnn', 5: 'dropout', 6: 'groupnorm', 7: 'linear', 8: 'addition', 9: 'identity'}


def _dev_index(device) -> int:
    if isinstance(device, int):
        return device
    d = torch.device(device)
    if d.type != 'cuda':
        raise _lib.EngineError(f'kraken_b200 runs on CUDA (sm_100a) devices only, got device {device!r}; there is no CPU path')
    return d.index if d.index is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)


def _as_f32(x) -> Union[torch.Tensor, np.ndarray]:
    if isinstance(x, torch.Tensor):
        return x.detach().to(torch.float32).contiguous()
    return np.ascontiguousarray(x, dtype=np.float32)


def _ptr(t) -> int:
    return t.data_ptr() if isinstance(t, torch.Tensor) else t.ctypes.data


def _on_device(t) -> bool:
    return isinstance(t, torch.Tensor) and t.is_cuda


def _stream_for(t, device: Optional[int] = None) -> Optional[int]:
    """The caller's current CUDA stream (thread-local in torch) - also for host inputs, so that several handles can be driven
    concurrently from different threads/streams."""
    if _on_device(t):
        return torch.cuda.current_stream(t.device).cuda_stream
    if device is not None and torch.cuda.is_available():
        return torch.cuda.current_stream(device).cuda_stream
    return None


class EngineNet:
    """The `nn` callable: `(x[N,C,H,W], seq_lens[N]|None) -> (y[N,C',H',W'], seq_lens'|None)`
    (contract of MultiParamSequential.forward, kraken/lib/vgsl/layers.py:44-53)."""

    def __init__(self, owner: 'TorchVGSLModel'):
        self._o = owner

    def __call__(self, x, seq_lens=None, output_shape=None):
        o = self._o
        o._ensure_finalized(x)
        x = _as_f32(x)
        if x.ndim != 4:
            raise ValueError(f'expected a 4D NCHW input, got shape {tuple(x.shape)}')
        n, c, h, w = (int(v) for v in x.shape)
        if c != o.input[1]:
            raise ValueError(f'expected {o.input[1]} input channels, got {c}')
        widths = None
        if seq_lens is not None:
            widths = np.ascontiguousarray(torch.as_tensor(seq_lens).cpu().numpy() if isinstance(seq_lens, torch.Tensor) else seq_lens, dtype=np.int32)
            if widths.shape != (n,):
                raise ValueError('seq_lens must have one entry per batch element')
        dims = (C.c_int32 * 4)()
        check(lib.kb_model_infer_dims(o._h, n, h, w, dims))
        on_dev = _on_device(x)
        if on_dev:
            if x.device.index != o._device:
                x = x.to(f'cuda:{o._device}')
            out = torch.empty(tuple(dims), dtype=torch.float32, device=x.device)
        else:
            out = torch.empty(tuple(dims), dtype=torch.float32)
        olens = np.zeros(n, dtype=np.int32)
        check(lib.kb_forward(o._h, _ptr(x), int(on_dev), n, h, w, widths.ctypes.data if widths is not None else None,
                             out.data_ptr(), int(on_dev), olens.ctypes.data, _stream_for(x, o._device)))
        return out, (torch.from_numpy(olens.astype(np.int64)) if seq_lens is not None else None)

    # state-dict style access used by tests / tooling
    def layer_output(self, name: str) -> torch.Tensor:
        o = self._o
        dims = (C.c_int32 * 4)()
        check(lib.kb_debug_layer_output(o._h, name.encode(), dims, None, 1))
        out = np.empty(tuple(dims), dtype=np.float32)
        check(lib.kb_debug_layer_output(o._h, name.encode(), dims, out.ctypes.data, 0))
        return torch.from_numpy(out)


class TorchVGSLModel:
    _kraken_min_version = '5.0.0'

    def __init__(self, **kwargs) -> None:
        self.user_metadata: dict[str, Any] = {}
        if (vgsl := kwargs.pop('vgsl', None)) is None:
            raise ValueError('vgsl specification argument is missing in args.')
        self._h = None
        self._device: Optional[int] = None
        self._finalized = False
        self._weights: dict[str, np.ndarray] = {}
        self.criterion = None
        self.codec = None
        self._build(vgsl)
        codec = kwargs.get('codec', None)
        if codec is not None:
            self.add_codec(codec if isinstance(codec, PytorchCodec) else PytorchCodec(codec))
        md = {'accuracy': [], 'metrics': [], 'seg_type': None, 'one_channel_mode': None, 'model_type': []}
        md.update(self.user_metadata)
        md.update(**{k: v for k, v in kwargs.items() if k != 'codec' or not isinstance(v, PytorchCodec)})
        self.user_metadata = md
        self.user_metadata['vgsl'] = '[' + ' '.join(self.named_spec) + ']'
        self.nn = EngineNet(self)
        self.init_weights()

    # ---- graph -----------------------------------------------------------------------------
    def _build(self, vgsl: str):
        h = C.c_void_p()
        check(lib.kb_model_create(vgsl.encode('utf-8'), C.byref(h)))
        if self._h is not None:
            lib.kb_model_destroy(self._h)
        self._h = h
        self._finalized = False
        self.spec = vgsl
        n = lib.kb_model_named_spec(self._h, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib.kb_model_named_spec(self._h, buf, n + 1)
        self._named = buf.value.decode('utf-8')
        self.named_spec = self._named[1:-1].split(' ')
        shp = (C.c_int32 * 4)()
        check(lib.kb_model_input_shape(self._h, shp))
        self.input = tuple(shp)
        check(lib.kb_model_output_shape(self._h, shp))
        self.output = tuple(shp)
        self._tensors: list[tuple[str, tuple[int, ...]]] = []
        name = C.create_string_buffer(512)
        shape = (C.c_int64 * 4)()
        nd = C.c_int32()
        for i in range(lib.kb_model_num_tensors(self._h)):
            check(lib.kb_model_tensor_info(self._h, i, name, 512, shape, C.byref(nd)))
            self._tensors.append((name.value.decode(), tuple(shape[:nd.value])))
        self.layers = []
        info = _lib.LayerInfo()
        for i in range(lib.kb_model_num_layers(self._h)):
            check(lib.kb_model_layer_info(self._h, i, C.byref(info)))
            self.layers.append({'kind': _KIND.get(info.kind, '?'), 'name': info.name.decode(), 'path': info.path.decode(),
                                'block': info.block.decode(), 'out_shape': tuple(info.out_shape)})

    def __del__(self):
        try:
            if getattr(self, '_h', None) is not None:
                lib.kb_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- weights ---------------------------------------------------------------------------
    def init_weights(self, idx: slice = slice(0, None)) -> None:
        """Same distributions as the reference (model.py:450-479): conv U(-0.1, 0.1), LSTM orthogonal
        with forget-gate bias 1, linear Xavier-uniform / zero bias, GroupNorm 1 / 0."""
        sel = set(l['path'] for l in self.layers[idx]) if idx != slice(0, None) else None
        for name, shape in self._tensors:
            path = name[3:].rsplit('.', 2)[0]
            if sel is not None and path not in sel:
                continue
            t = torch.empty(shape)
            if '.co.' in name:
                torch.nn.init.uniform_(t, -0.1, 0.1)
            elif name.endswith('.lin.weight'):
                torch.nn.init.xavier_uniform_(t)
            elif name.endswith('.lin.bias'):
                t.zero_()
            elif '.layer.weight_' in name and len(shape) == 1:
                torch.nn.init.uniform_(t, -0.1, 0.1)       # peephole vectors of the ocropy cell (the reference leaves them uninitialised)
            elif '.layer.weight_' in name:
                torch.nn.init.orthogonal_(t)
            elif '.layer.bias_' in name:
                k = 1.0 / np.sqrt(shape[0] // 4)
                torch.nn.init.uniform_(t, -k, k)
                t[len(t) // 4:len(t) // 2] = 1.0
            elif name.endswith('.layer.weight'):
                t.fill_(1.0)
            else:
                t.zero_()
            self._weights[name] = t.numpy().copy()
        self._finalized = False

    def state_dict(self) -> dict[str, torch.Tensor]:
        return {k: torch.from_numpy(self._weights[k].copy()) for k, _ in self._tensors}

    def load_state_dict(self, state_dict: dict, strict: bool = True):
        expected = dict(self._tensors)
        missing = [k for k in expected if k not in state_dict]
        unexpected = [k for k in state_dict if k not in expected]
        if strict and (missing or unexpected):
            raise RuntimeError(f'Error(s) in loading state_dict for TorchVGSLModel:\n    Missing key(s): {missing}\n    Unexpected key(s): {unexpected}')
        for k, shp in expected.items():
            if k not in state_dict:
                continue
            v = state_dict[k]
            a = v.detach().cpu().float().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, dtype=np.float32)
            if tuple(a.shape) != tuple(shp):
                raise RuntimeError(f'size mismatch for {k}: copying a param with shape {tuple(a.shape)}, the shape in current model is {tuple(shp)}.')
            self._weights[k] = np.ascontiguousarray(a, dtype=np.float32)
        self._finalized = False
        return missing, unexpected

    def _push_weights(self):
        for k, shp in self._tensors:
            a = self._weights[k]
            shape = (C.c_int64 * len(shp))(*shp)
            check(lib.kb_model_load_tensor(self._h, k.encode(), a.ctypes.data, shape, len(shp)))

    def to(self, device):
        self._device = _dev_index(device)
        self._push_weights()
        check(lib.kb_model_finalize(self._h, self._device))
        self._finalized = True
        return self

    def cuda(self, device=0):
        return self.to(f'cuda:{device}' if isinstance(device, int) else device)

    def eval(self):
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError('kraken_b200 is an inference engine; training is out of scope')
        return self

    def _ensure_finalized(self, x=None):
        if self._finalized:
            return
        if self._device is None:
            self._device = x.device.index if _on_device(x) else 0
        self.to(self._device)

    # ---- reference surface ----------------------------------------------------------------
    def forward(self, x, seq_lens=None):
        return self.nn(x, seq_lens)

    __call__ = forward

    def add_codec(self, codec: PytorchCodec) -> None:
        self.codec = codec
        self.user_metadata['codec'] = json.dumps(self.codec.c2l)

    @property
    def one_channel_mode(self):
        return self.user_metadata['one_channel_mode']

    @one_channel_mode.setter
    def one_channel_mode(self, val):
        if val not in ['1', 'L', None]:
            raise ValueError('one_channel_mode {} is not one of [1, L, None]'.format(val))
        self.user_metadata['one_channel_mode'] = val

    @property
    def model_type(self):
        return self.user_metadata.get('model_type', [])

    @model_type.setter
    def model_type(self, val):
        if isinstance(val, str):
            val = [val]
        for v in val:
            if v not in ['recognition', 'segmentation']:
                raise ValueError('model_type {} is not one of [recognition, segmentation]'.format(v))
        self.user_metadata['model_type'] = val

    @property
    def seg_type(self):
        return self.user_metadata.get('seg_type', None)

    @seg_type.setter
    def seg_type(self, val):
        if val not in ['bbox', 'baselines', None]:
            raise ValueError('segmentation type {} is not one of [bbox, baselines, None]'.format(val))
        self.user_metadata['seg_type'] = val

    @property
    def hyper_params(self):
        return self.user_metadata['hyper_params']

    @hyper_params.setter
    def hyper_params(self, val):
        self.user_metadata.setdefault('hyper_params', {}).update(val)

    @property
    def aux_layers(self):
        return {}

    @property
    def use_legacy_polygons(self):
        return self.user_metadata.get('legacy_polygons', True)

    @use_legacy_polygons.setter
    def use_legacy_polygons(self, val: bool):
        self.user_metadata['legacy_polygons'] = val

    # ---- spec surgery (model.py:245-268, 548-568) ------------------------------------------
    def resize_output(self, output_size: int, del_indices: Optional[Iterable] = None) -> None:
        last = self.layers[-1]
        if last['kind'] not in ('conv', 'linear'):
            raise ValueError('last layer is neither linear nor convolutional layer')
        m = re.match(r'(O)(?P<name>{\w+})?(?P<dim>2|1|0)(?P<type>l|s|c)(?P<aug>a)?(?P<out>\d+)', self.named_spec[-1])
        if not m:
            raise ValueError('Output specification is not parsable')
        del_indices = sorted(set(del_indices or []))
        pre = 'nn.' + last['path']
        wk, bk = (pre + '.co.weight', pre + '.co.bias') if last['kind'] == 'conv' else (pre + '.lin.weight', pre + '.lin.bias')
        old_w, old_b = self._weights[wk], self._weights[bk]
        keep = [i for i in range(old_w.shape[0]) if i not in del_indices]
        named = list(self.named_spec)
        named[-1] = 'O{}{}{}{}{}'.format(m.group('name'), m.group('dim'), m.group('type'), m.group('aug') or '', output_size)
        saved = dict(self._weights)
        md = self.user_metadata
        self._build('[' + ' '.join(named) + ']')
        self._weights = saved
        self.nn = EngineNet(self)
        # new rows initialised like a fresh layer, surviving rows copied (layers.py resize())
        new_w = torch.empty((output_size,) + old_w.shape[1:])
        if last['kind'] == 'conv':
            torch.nn.init.uniform_(new_w, -0.1, 0.1)
            new_b = torch.empty(output_size).uniform_(-0.1, 0.1)
        else:
            torch.nn.init.xavier_uniform_(new_w)
            new_b = torch.zeros(output_size)
        k = min(len(keep), output_size)
        new_w[:k] = torch.from_numpy(old_w[keep[:k]])
        new_b[:k] = torch.from_numpy(old_b[keep[:k]])
        self._weights[wk], self._weights[bk] = new_w.numpy().copy(), new_b.numpy().copy()
        self.user_metadata = md
        self.spec = '[' + ' '.join(self.named_spec) + ']'
        self.user_metadata['vgsl'] = self.spec

    def append(self, idx: int, spec: str) -> None:
        """Splits the model at layer `idx` and appends the layers in `spec` (freshly initialised)."""
        named = self.named_spec[:idx + 1] + spec.strip()[1:-1].split(' ')
        saved = dict(self._weights)
        md = self.user_metadata
        self._build('[' + ' '.join(named) + ']')
        self.nn = EngineNet(self)
        self._weights = {}
        self.init_weights()
        for k, shp in self._tensors:
            if k in saved and tuple(saved[k].shape) == tuple(shp) and any(k.startswith('nn.' + l['path'] + '.') for l in self.layers[:idx]):
                self._weights[k] = saved[k]
        self.user_metadata = md
        self.spec = '[' + ' '.join(self.named_spec) + ']'
        self.user_metadata['vgsl'] = self.spec

    # ---- files -----------------------------------------------------------------------------
    @classmethod
    def load_model(cls, path: str):
        """Loads a safetensors or CoreML kraken model file (first VGSL model in it)."""
        from .weights import load_model_file
        files = load_model_file(str(path))
        if not files:
            raise ValueError(f'No VGSL model found in {path}')
        return cls.from_model_file(files[0])

    @classmethod
    def from_model_file(cls, mf):
        md = dict(mf.metadata)
        m = cls(vgsl=mf.vgsl, codec=mf.codec, **md)
        m.load_state_dict(mf.weights)
        return m

    # ---- inference surface (model.py:491-546) -------------------------------------------------
    def prepare_for_inference(self, config=None):
        self._inf_config = config
        dev = getattr(config, 'device', None) if config is not None else None
        if dev is None or dev == 'auto' or (isinstance(dev, str) and dev == 'cpu'):
            dev = 0
        self.to(dev if not isinstance(dev, (list, tuple)) else dev[0])
        return self

    @torch.inference_mode()
    def predict(self, *args, **kwargs):
        if 'recognition' in self.model_type:
            from .rpred import recognize_lines
            return recognize_lines(self, *args, **kwargs)
        elif 'segmentation' in self.model_type:
            from .blla import compute_segmentation_map
            return compute_segmentation_map(self, *args, **kwargs)
        raise ValueError(f'{self} has no model_type set')

    # ---- engine extras -----------------------------------------------------------------------
    def infer_dims(self, n: int, h: int, w: int):
        dims = (C.c_int32 * 4)()
        check(lib.kb_model_infer_dims(self._h, n, h, w, dims))
        return tuple(dims)

    def infer_lens(self, h: int, w: int, widths: Sequence[int]):
        widths = np.ascontiguousarray(widths, dtype=np.int32)
        out = np.zeros_like(widths)
        check(lib.kb_model_infer_lens(self._h, len(widths), h, w, widths.ctypes.data, out.ctypes.data))
        return out

    @property
    def launch_count(self) -> int:
        return int(lib.kb_launch_count(self._h))

    @property
    def range_fallback_count(self) -> int:
        """Calls that were repeated on the fp32 CUDA-core kernels because an activation left the fp16 operand range."""
        return int(lib.kb_range_fallback_count(self._h))

    def reset_launch_count(self):
        lib.kb_reset_launch_count(self._h)

    def set_timing(self, on: bool):
        check(lib.kb_set_timing(self._h, int(on)))

    def last_timing(self) -> list[tuple[str, float]]:
        """[(stage name, device ms)] of the most recent call, in execution order (needs set_timing(True))."""
        out = []
        name = C.create_string_buffer(128)
        ms = C.c_float()
        for i in range(max(0, lib.kb_timing_count(self._h))):
            check(lib.kb_timing_entry(self._h, i, name, 128, C.byref(ms)))
            out.append((name.value.decode(), float(ms.value)))
        return out
