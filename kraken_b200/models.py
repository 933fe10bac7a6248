"""
Mirror of kraken.lib.models.TorchSeqRecognizer / load_any (reference: kraken/lib/models.py:31-199) on top of
the engine.  Duck-type compatible with what `kraken.rpred.mm_rpred` touches: `.nn` (with `.input`,
`.one_channel_mode`, `.use_legacy_polygons`), `.codec`, `.seg_type`, `.predict()`, `.outputs`.

With the default decoder, `predict*` run the fused device path (`kb_recognize`: net -> softmax statistics ->
arg-max -> CTC collapse in one call, only the label tuples come back).  A user-supplied `decoder` gets the
reference behaviour: probabilities as a (N, C, W) numpy array, then the hook.
"""
from __future__ import annotations

import ctypes as C
from os.path import abspath, expanduser, expandvars
from typing import Optional, Union

import numpy as np
import torch

from . import ctc_decoder
from ._lib import KrakenInputException, check, lib
from .vgsl import TorchVGSLModel, _as_f32, _on_device, _ptr, _stream_for

__all__ = ['TorchSeqRecognizer', 'load_any', 'KrakenInvalidModelException']


class KrakenInvalidModelException(Exception):
    pass


class _LazyOutputs:
    """Stands in for `TorchSeqRecognizer.outputs` (the (N, C, W) probability array of kraken/lib/models.py:116) after a fused call
    that only fetched label blocks: it knows its shape - all `kraken.rpred.mm_rpred` reads (rpred.py:229,299) - and computes the
    probabilities only if somebody actually looks at them."""

    def __init__(self, rec, line, lens, shape):
        self._rec, self._line, self._lens, self.shape, self._arr = rec, line, lens, tuple(shape), None
        self.ndim = 3

    def _get(self) -> np.ndarray:
        if self._arr is None:
            rec, self._rec = self._rec, None
            keep = rec.outputs
            rec._recognize_raw(self._line, self._lens, want_probs=True)
            self._arr, self._line, self._lens = rec.outputs, None, None
            if keep is not self:
                rec.outputs = keep
        return self._arr

    def __array__(self, dtype=None, copy=None):
        a = self._get()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, idx):
        return self._get()[idx]

    def __len__(self):
        return self.shape[0]


class TorchSeqRecognizer:
    def __init__(self, nn: TorchVGSLModel, decoder=ctc_decoder.greedy_decoder, temperature: float = 1.0,
                 train: bool = False, device: str = 'cuda:0'):
        if train:
            raise NotImplementedError('kraken_b200 is an inference engine; train=True is not supported')
        self.nn = nn
        self.kind = ''
        self.codec = self.nn.codec
        self.decoder = decoder
        self.temperature = temperature
        self.train = train
        self.device = device
        if nn.model_type and 'recognition' not in nn.model_type:
            raise ValueError(f'Models of type {nn.model_type} are not supported by TorchSeqRecognizer')
        self.one_channel_mode = nn.one_channel_mode
        self.seg_type = nn.seg_type
        self.outputs = None
        self.keep_outputs = False      # True: predict* also fetch the (N, C, W) probabilities into `self.outputs` (models.py:116)
        self._dims_cache = {}          # (engine handle, n, h, w) -> output dims of nn
        if self.device:
            self.nn.to(device)

    def to(self, device):
        self.device = device
        self.nn.to(device)

    # -- fused device path ----------------------------------------------------------------------
    def _recognize(self, line, lens, want_probs: bool):
        r = self._recognize_raw(line, lens, want_probs)
        return ctc_decoder.unpack_decoded(r['labels'], r['starts'], r['ends'], r['confs'], r['counts']), r['olens']

    def _recognize_raw(self, line, lens, want_probs: bool, out: Optional[dict] = None) -> dict:
        """One `kb_recognize` call; returns the ABI's fixed-stride output blocks as numpy arrays
        (labels/starts/ends/confs [N, T], counts [N], olens [N] or None).  `out` may supply preallocated C-contiguous
        blocks (e.g. views into one pinned buffer that is shipped elsewhere afterwards) for labels/starts/ends/confs/counts."""
        net = self.nn
        net._ensure_finalized(line)
        x = _as_f32(line)
        if x.ndim != 4:
            raise ValueError(f'expected a 4D NCHW input, got shape {tuple(x.shape)}')
        n, c, h, w = (int(v) for v in x.shape)
        if c != net.input[1]:
            raise ValueError(f'expected {net.input[1]} input channels, got {c}')
        if _on_device(x) and x.device.index != net._device:
            x = x.to(f'cuda:{net._device}')
        widths = None
        if lens is not None:
            widths = np.ascontiguousarray(torch.as_tensor(lens).cpu().numpy(), dtype=np.int32)
            if widths.shape != (n,):
                raise ValueError('seq_lens must have one entry per batch element')
        key = (getattr(net._h, 'value', None), getattr(net, 'spec', None), n, h, w)
        dims = self._dims_cache.get(key)
        if dims is None:
            dims = self._dims_cache[key] = net.infer_dims(n, h, w)
        if dims[2] != 1:
            raise KrakenInputException('Expected dimension 3 to be 1, actual {}'.format(tuple(dims)))
        T, ncls = dims[3], dims[1]
        stride = max(T, 1)
        # the engine writes every element of these blocks (valid prefix + zero fill), so they need no initialisation
        if out is not None:
            labels, starts, ends, confs, counts = out['labels'], out['starts'], out['ends'], out['confs'], out['counts']
            for a, dt, shp in ((labels, np.int32, (n, stride)), (starts, np.int32, (n, stride)), (ends, np.int32, (n, stride)),
                               (confs, np.float32, (n, stride)), (counts, np.int32, (n,))):
                if a.dtype != dt or a.shape != shp or not a.flags['C_CONTIGUOUS']:
                    raise ValueError('preallocated output block has the wrong dtype, shape or layout')
        else:
            labels = np.empty((n, stride), np.int32)
            starts = np.empty((n, stride), np.int32)
            ends = np.empty((n, stride), np.int32)
            confs = np.empty((n, stride), np.float32)
            counts = np.empty(n, np.int32)
        olens = np.zeros(n, np.int32)
        probs = np.empty((n, ncls, T), np.float32) if want_probs else None
        on_dev = _on_device(x)
        check(lib.kb_recognize(net._h, _ptr(x), int(on_dev), n, h, w, widths.ctypes.data if widths is not None else None,
                               float(self.temperature), labels.ctypes.data, starts.ctypes.data, ends.ctypes.data, confs.ctypes.data,
                               counts.ctypes.data, stride, olens.ctypes.data, probs.ctypes.data if probs is not None else None, 0,
                               _stream_for(x, net._device)))
        self.outputs = probs if probs is not None else _LazyOutputs(self, line, lens, (n, ncls, T))
        return {'labels': labels, 'starts': starts, 'ends': ends, 'confs': confs, 'counts': counts,
                'olens': olens if lens is not None else None}

    def recognize_u8(self, lines_u8, widths=None, invert_max=None) -> dict:
        """One `kb_recognize_u8` call on uint8 line images (N, C, H, W) as `v2.PILToTensor()` yields them; scaling to
        [0, 1], `tensor_invert` and the zero right-padding of `ImageInputTransforms` / `_recognize_*_lines`
        (kraken/lib/dataset/utils.py:148-151, kraken/lib/vgsl/rpred.py:129-131) run on the device, bit-identical to the
        reference.  `invert_max`: per line `int(im.max())` of the un-padded crop (None: no inversion).  Returns the same
        blocks as `_recognize_raw`."""
        net = self.nn
        x = lines_u8 if isinstance(lines_u8, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(lines_u8))
        if x.dtype != torch.uint8 or x.ndim != 4:
            raise ValueError('expected a uint8 NCHW tensor')
        x = x.contiguous()
        net._ensure_finalized(x)
        n, c, h, w = (int(v) for v in x.shape)
        if c != net.input[1]:
            raise ValueError(f'expected {net.input[1]} input channels, got {c}')
        if x.is_cuda and x.device.index != net._device:
            x = x.to(f'cuda:{net._device}')
        wd = np.ascontiguousarray(torch.as_tensor(widths).cpu().numpy(), dtype=np.int32) if widths is not None else None
        inv = np.ascontiguousarray(np.asarray(invert_max), dtype=np.int16) if invert_max is not None else None
        dims = net.infer_dims(n, h, w)
        if dims[2] != 1:
            raise KrakenInputException('Expected dimension 3 to be 1, actual {}'.format(tuple(dims)))
        T = dims[3]
        stride = max(T, 1)
        labels = np.zeros((n, stride), np.int32); starts = np.zeros((n, stride), np.int32); ends = np.zeros((n, stride), np.int32)
        confs = np.zeros((n, stride), np.float32); counts = np.zeros(n, np.int32); olens = np.zeros(n, np.int32)
        check(lib.kb_recognize_u8(net._h, x.data_ptr(), int(x.is_cuda), n, h, w, wd.ctypes.data if wd is not None else None,
                                  inv.ctypes.data if inv is not None else None, float(self.temperature), labels.ctypes.data,
                                  starts.ctypes.data, ends.ctypes.data, confs.ctypes.data, counts.ctypes.data, stride, olens.ctypes.data,
                                  None, 0, _stream_for(x, net._device)))
        return {'labels': labels, 'starts': starts, 'ends': ends, 'confs': confs, 'counts': counts,
                'olens': olens if widths is not None else None}

    # -- record assembly on the device -----------------------------------------------
    def recognize_records(self, line, lens, orig_widths, padding: int = 16, invert_max=None):
        """`_recognize_*_lines` up to the record fields (kraken/lib/vgsl/rpred.py:126-157): one engine call returns, per line,
        (text, [[start, end]] in the coordinates of the ORIGINAL line image - `_scale_val`, rpred.py:231 - and confidences).
        Code-point lookup and position scaling run inside the CTC collapse kernel (`kb_recognize_records`); needs a 1:1 codec
        (one label - one code point, the common case; anything else: `predict` + the reference's own record code).
        `line`: float32 (N, C, H, W) as `predict` takes it, or uint8 as `recognize_u8` (then `invert_max` applies)."""
        net = self.nn
        if self.codec is None or self.codec._single_lut() is None:
            raise ValueError('recognize_records needs a codec in which every code is one label and one code point')
        is_u8 = isinstance(line, torch.Tensor) and line.dtype == torch.uint8 or isinstance(line, np.ndarray) and line.dtype == np.uint8
        x = (line if isinstance(line, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(line))).contiguous() if is_u8 else _as_f32(line)
        net._ensure_finalized(x)
        n, c, h, w = (int(v) for v in x.shape)
        if c != net.input[1]:
            raise ValueError(f'expected {net.input[1]} input channels, got {c}')
        if _on_device(x) and x.device.index != net._device:
            x = x.to(f'cuda:{net._device}')
        if getattr(self, '_codec_pushed', None) is not self.codec:
            lut = self.codec._single_lut()
            table = np.zeros(lut.shape[0], np.uint32)
            for i, ch in enumerate(lut.tolist()):
                table[i] = ord(ch) if ch else 0
            check(lib.kb_model_set_codec(net._h, table.ctypes.data, int(table.shape[0])))
            self._codec_pushed = self.codec
        widths = np.ascontiguousarray(torch.as_tensor(lens).cpu().numpy(), dtype=np.int32) if lens is not None else np.full(n, w, np.int32)
        ow = np.ascontiguousarray(np.asarray(orig_widths), dtype=np.int32)
        if widths.shape != (n,) or ow.shape != (n,):
            raise ValueError('seq_lens and orig_widths must have one entry per batch element')
        inv = np.ascontiguousarray(np.asarray(invert_max), dtype=np.int16) if invert_max is not None else None
        dims = net.infer_dims(n, h, w)
        if dims[2] != 1:
            raise KrakenInputException('Expected dimension 3 to be 1, actual {}'.format(tuple(dims)))
        stride = max(dims[3], 1)
        cps = np.empty((n, stride), np.uint32); starts = np.empty((n, stride), np.int32); ends = np.empty((n, stride), np.int32)
        confs = np.empty((n, stride), np.float32); counts = np.empty(n, np.int32); olens = np.zeros(n, np.int32)
        on_dev = _on_device(x)
        check(lib.kb_recognize_records(net._h, _ptr(x), 1 if is_u8 else 0, int(on_dev), n, h, w, widths.ctypes.data,
                                       inv.ctypes.data if inv is not None else None, float(self.temperature), ow.ctypes.data, int(padding),
                                       cps.ctypes.data, starts.ctypes.data, ends.ctypes.data, confs.ctypes.data, counts.ctypes.data, stride,
                                       olens.ctypes.data, _stream_for(x, net._device)))
        out = []
        for i in range(n):
            k = int(min(counts[i], stride))
            keep = cps[i, :k] != 0
            if self.codec.strict and not keep.all():
                from .codec import KrakenEncodeException
                raise KrakenEncodeException('Non-decodable sequence encountered.')
            text = ''.join(map(chr, cps[i, :k][keep].tolist()))
            out.append((text, np.stack([starts[i, :k][keep], ends[i, :k][keep]], 1).tolist(), confs[i, :k][keep].tolist()))
        return out

    # -- asynchronous pipeline (kb_recognize_async / kb_wait) --------------------------------------------
    def set_pipeline_depth(self, depth: int) -> None:
        """Number of batches one host thread can keep in flight on this model (own stream + workspace per slot, one copy of
        the weights)."""
        self.nn._ensure_finalized(None)
        check(lib.kb_set_pipeline_depth(self.nn._h, int(depth)))
        self._depth = int(depth)

    def submit(self, line, lens=None, invert_max=None) -> int:
        """Enqueues one batch - float32 lines (N, C, H, W) as `predict` takes them, or the uint8 lines `recognize_u8` takes - and
        returns a ticket without waiting for the GPU.  Host tensors should be pinned (`tensor.pin_memory()`); they are kept alive
        until `collect`."""
        net = self.nn
        if not getattr(self, '_depth', 0):
            self.set_pipeline_depth(4)
        is_u8 = isinstance(line, torch.Tensor) and line.dtype == torch.uint8 or isinstance(line, np.ndarray) and line.dtype == np.uint8
        if is_u8:
            x = line if isinstance(line, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(line))
            x = x.contiguous()
        else:
            x = _as_f32(line)
        if x.ndim != 4:
            raise ValueError(f'expected a 4D NCHW input, got shape {tuple(x.shape)}')
        n, c, h, w = (int(v) for v in x.shape)
        if c != net.input[1]:
            raise ValueError(f'expected {net.input[1]} input channels, got {c}')
        if _on_device(x) and x.device.index != net._device:
            x = x.to(f'cuda:{net._device}')
        widths = np.ascontiguousarray(torch.as_tensor(lens).cpu().numpy(), dtype=np.int32) if lens is not None else None
        if widths is not None and widths.shape != (n,):
            raise ValueError('seq_lens must have one entry per batch element')
        inv = np.ascontiguousarray(np.asarray(invert_max), dtype=np.int16) if invert_max is not None else None
        key = (getattr(net._h, 'value', None), getattr(net, 'spec', None), n, h, w)
        dims = self._dims_cache.get(key)
        if dims is None:
            dims = self._dims_cache[key] = net.infer_dims(n, h, w)
        if dims[2] != 1:
            raise KrakenInputException('Expected dimension 3 to be 1, actual {}'.format(tuple(dims)))
        stride = max(dims[3], 1)
        t = C.c_int64()
        on_dev = _on_device(x)
        check(lib.kb_recognize_async(net._h, _ptr(x), 1 if is_u8 else 0, int(on_dev), n, h, w,
                                     widths.ctypes.data if widths is not None else None, inv.ctypes.data if inv is not None else None,
                                     float(self.temperature), stride, _stream_for(x, net._device) if on_dev else None, C.byref(t)))
        if not hasattr(self, '_inflight'):
            self._inflight = {}
        self._inflight[t.value] = (x, n, stride, lens is not None)
        return t.value

    def collect(self, ticket: int, out: Optional[dict] = None) -> dict:
        """Waits for a submitted batch (the thread sleeps on a blocking CUDA event) and returns the same blocks as `_recognize_raw`."""
        x, n, stride, has_lens = self._inflight.pop(ticket)
        if out is not None:
            labels, starts, ends, confs, counts = out['labels'], out['starts'], out['ends'], out['confs'], out['counts']
        else:
            labels = np.empty((n, stride), np.int32); starts = np.empty((n, stride), np.int32); ends = np.empty((n, stride), np.int32)
            confs = np.empty((n, stride), np.float32); counts = np.empty(n, np.int32)
        olens = np.zeros(n, np.int32)
        check(lib.kb_wait(self.nn._h, int(ticket), labels.ctypes.data, starts.ctypes.data, ends.ctypes.data, confs.ctypes.data,
                          counts.ctypes.data, olens.ctypes.data))
        return {'labels': labels, 'starts': starts, 'ends': ends, 'confs': confs, 'counts': counts, 'olens': olens if has_lens else None}

    def recognize_stream(self, batches, depth: int = 4):
        """Generator: feeds `(lines, lens)` pairs (or `(uint8 lines, lens, invert_max)` triples) through the engine with `depth` batches
        in flight from this one thread and yields their result blocks in order."""
        if getattr(self, '_depth', 0) != depth:
            self.set_pipeline_depth(depth)
        pending = []
        for b in batches:
            if len(pending) == depth:
                yield self.collect(pending.pop(0))
            pending.append(self.submit(*b))
        while pending:
            yield self.collect(pending.pop(0))

    # -- reference surface ------------------------------------------------------------------------
    def forward(self, line: torch.Tensor, lens: Optional[torch.Tensor] = None):
        """(N, C, H, W) lines -> ((N, C, W) softmax numpy array, output lengths) - models.py:93-119."""
        r = self._recognize_raw(line, lens, want_probs=True)
        return self.outputs, r['olens']

    def _decode(self, line, lens):
        if lens is None and getattr(line, 'ndim', 0) == 4 and int(line.shape[0]) > 1:
            raise ValueError('seq_lens need to be set for batch decoding.')          # ctc_decoder.py:60-61
        if self.decoder is ctc_decoder.greedy_decoder:
            # fused device path: only the label blocks come back.  `return_logits` (the reference's RecognitionInferenceConfig flag,
            # rpred.py:227) or the legacy callers that read `self.outputs` set `keep_outputs`.
            dec, _ = self._recognize(line, lens, want_probs=bool(self.keep_outputs))
            return dec
        o, olens = self.forward(line, lens)
        return self.decoder(o, olens)

    def predict(self, line, lens=None) -> list[list[tuple[str, int, int, float]]]:
        return [self.codec.decode(locs) for locs in self._decode(line, lens)]

    def predict_string(self, line, lens=None) -> list[str]:
        return [''.join(x[0] for x in self.codec.decode(locs)) for locs in self._decode(line, lens)]

    def predict_labels(self, line, lens=None) -> list[list[tuple[int, int, int, float]]]:
        return self._decode(line, lens)

    def predict_records(self, line, lens=None) -> list[tuple[str, np.ndarray, np.ndarray, np.ndarray]]:
        """Throughput variant of `predict`: per line (text, starts, ends, confidences) assembled from the
        engine's output blocks with one table lookup per line (`PytorchCodec.decode_blocks`) instead of per-character tuples.
        Same characters, positions and confidences as `predict`."""
        r = self._recognize_raw(line, lens, want_probs=False)
        return self.codec.decode_blocks(r['labels'], r['starts'], r['ends'], r['confs'], r['counts'])


def load_any(fname: Union[str, 'object'], train: bool = False, device: str = 'cuda:0') -> TorchSeqRecognizer:
    fname = abspath(expandvars(expanduser(str(fname))))
    try:
        nn = TorchVGSLModel.load_model(fname)
    except Exception as e:
        raise KrakenInvalidModelException('File {} not loadable by any parser.'.format(fname)) from e
    seq = TorchSeqRecognizer(nn, train=train, device=device)
    seq.kind = 'vgsl'
    return seq
