"""
Reference-side binding: `accelerate(model)` turns a loaded **kraken** model (the reference's own
`kraken.lib.vgsl.TorchVGSLModel`, an `nn.Module`) into one whose hot path runs on the engine, in place - the object keeps its
class, parameters, `state_dict`, codec and metadata, so `isinstance` checks (CoreML writer `kraken/models/writers.py:101`, forced
alignment) and serialisation keep working.  What is rebound (reference file:line):

  model._rec_predict(line, lens)        kraken/lib/vgsl/rpred.py:210-229   -> one fused engine call (net -> softmax statistics ->
                                        arg-max -> CTC collapse); honours `return_logits` (rpred.py:200,227: `self.outputs` is then the
                                        (N, C, W) probability tensor) and a custom `decoder` hook (kraken/configs/base.py:235)
  model._compute_segmentation_map(im)   kraken/lib/vgsl/spred.py:237-287   -> nn + nearest upsample + sigmoid in `kb_segment`
  model.nn.forward(x, seq_lens)         kraken/lib/vgsl/layers.py:44-53    -> `kb_forward` (legacy `TorchSeqRecognizer.forward`,
                                        `kraken.blla.compute_segmentation_map`, anything else that calls `model.nn(...)`)

`load_accelerated(path)` is the `kraken.loaders` entry point (pyproject.toml): it reads the file with kraken's own loaders and
accelerates every VGSL model in it when a Blackwell GPU is present (raising ValueError = "not mine" otherwise, which makes
`kraken.models.load_models` fall through to the stock loaders, loaders.py:35-43).  `TorchVGSLModelB200` is the additional
`kraken.models` registry name (names must be unique, kraken/models/utils.py:20-23): a subclass of the reference class that
accelerates itself in `prepare_for_inference`.

Nothing here imports kraken at module import time: the package stays importable without it.
"""
from __future__ import annotations

import types
from typing import Optional

import numpy as np
import torch

from . import ctc_decoder
from ._lib import lib
from .models import TorchSeqRecognizer
from .vgsl import TorchVGSLModel as EngineModel

__all__ = ['accelerate', 'load_accelerated', 'make_registry_class']


# ---- the three engine calls (module-level so that tests can substitute them) --------------------------------------------------------
def _engine_recognize(rec: TorchSeqRecognizer, line, lens, want_probs: bool) -> dict:
    return rec._recognize_raw(line, lens, want_probs=want_probs)


def _engine_forward(net: EngineModel, x, seq_lens):
    return net.nn(x, seq_lens)


def _engine_segment(net: EngineModel, pages, size):
    from .blla import segmentation_heatmap
    return segmentation_heatmap(net, pages, size)


def _twin(model, device) -> EngineModel:
    """Engine model with the reference model's graph and weights (spec: user_metadata['vgsl'], model.py:198-199)."""
    spec = model.user_metadata.get('vgsl') or model.spec
    md = {k: v for k, v in model.user_metadata.items() if k not in ('vgsl', 'codec')}
    twin = EngineModel(vgsl=spec, **md)
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items() if k.startswith('nn.')}
    twin.load_state_dict(sd)
    if device is not None:
        d = torch.device(device)
        twin._device = d.index if d.index is not None else 0          # finalised lazily on the first call (needs the GPU)
    return twin


def _is_default_decoder(dec) -> bool:
    if dec is None or dec is ctc_decoder.greedy_decoder:
        return True
    return getattr(dec, '__module__', '') == 'kraken.lib.ctc_decoder' and getattr(dec, '__name__', '') == 'greedy_decoder'


def accelerate(model, device: Optional[str] = 'cuda:0'):
    """Rebinds the hot path of a kraken `TorchVGSLModel` to the engine (see module docstring) and returns the same object.
    Idempotent; `model._b200` holds the engine twin.  Weights are snapshotted at call time: call again after changing them."""
    twin = _twin(model, device)
    rec = None
    if 'recognition' in (model.model_type or []) or getattr(model, 'codec', None) is not None:
        rec = TorchSeqRecognizer.__new__(TorchSeqRecognizer)
        rec.nn, rec.kind, rec.codec, rec.decoder, rec.temperature, rec.train, rec.device = twin, 'vgsl', model.codec, ctc_decoder.greedy_decoder, 1.0, False, device
        rec.one_channel_mode, rec.seg_type, rec.outputs, rec.keep_outputs, rec._dims_cache = model.one_channel_mode, model.seg_type, None, False, {}
    model._b200 = twin
    model._b200_rec = rec

    def nn_forward(self_nn, x, seq_lens=None, output_shape=None):
        out, olens = _engine_forward(twin, x, seq_lens)
        return out, olens

    # instance attribute on the existing MultiParamSequential: nn.Module.__call__ dispatches to it, parameters / state_dict /
    # indexing (`model.nn[-1]`) stay untouched
    model.nn.forward = types.MethodType(nn_forward, model.nn)

    def _rec_predict(self, line, lens=None):
        cfg = getattr(self, '_inf_config', None)
        temperature = float(getattr(cfg, 'temperature', 1.0))
        decoder = getattr(cfg, 'decoder', None)
        want_logits = bool(getattr(cfg, 'return_logits', False))
        rec.temperature = temperature
        fused = _is_default_decoder(decoder)
        r = _engine_recognize(rec, line, lens, want_probs=want_logits or not fused)
        n = int(line.shape[0])
        olens = torch.as_tensor(r['olens'].astype(np.int64)) if r['olens'] is not None else torch.full((n,), int(r['labels'].shape[1]), dtype=torch.long)
        if rec.outputs is not None and (want_logits or not fused):
            self.outputs = torch.from_numpy(np.asarray(rec.outputs))               # (N, C, W) like rpred.py:227
        if fused:
            decoded = ctc_decoder.unpack_decoded(r['labels'], r['starts'], r['ends'], r['confs'], r['counts'])
        else:
            decoded = decoder(self.outputs, olens)
        return [self.codec.decode(locs) for locs in decoded], olens

    if rec is not None:
        model._rec_predict = types.MethodType(_rec_predict, model)

    def _compute_segmentation_map(self, im):
        import torch.nn.functional as F  # noqa: F401  (kept for parity with the reference's imports)
        from kraken.lib.dataset import ImageInputTransforms
        from torchvision.transforms import v2
        batch, channels, height, width = self.input
        padding = self._inf_config.input_padding
        if isinstance(padding, int):
            padding = (padding,) * 4
        elif len(padding) == 2:
            padding = (padding[0], padding[0], padding[1], padding[1])
        transforms = ImageInputTransforms(batch, height, width, channels, padding, valid_norm=False, dtype=torch.float32)
        tf_idx, _ = next(filter(lambda x: isinstance(x[1], v2.PILToTensor), enumerate(transforms.transforms)))
        res_tf = v2.Compose(transforms.transforms[:tf_idx])
        scal_im = np.array(res_tf(im).convert('L'))
        tensor_im = transforms(im)
        o = _engine_segment(twin, tensor_im.unsqueeze(0), scal_im.shape)            # nn -> interpolate -> sigmoid (spred.py:268-272)
        pad = [p if p else None for p in padding]
        pad[1] = -pad[1] if pad[1] else None
        pad[3] = -pad[3] if pad[3] else None
        o = o[:, :, pad[2]:pad[3], pad[0]:pad[1]]
        scal_im = scal_im[pad[2]:pad[3], pad[0]:pad[1]]
        o = o.squeeze().cpu().float().numpy()
        scale = np.divide(im.size, o.shape[:0:-1])
        return {'heatmap': o, 'cls_map': self.user_metadata['class_mapping'],
                'bounding_regions': self.user_metadata.get('bounding_regions', None), 'scale': scale, 'scal_im': scal_im}

    if 'segmentation' in (model.model_type or []):
        model._compute_segmentation_map = types.MethodType(_compute_segmentation_map, model)
    return model


def _gpu_available() -> bool:
    return int(lib.kb_device_count()) > 0


def load_accelerated(path, tasks=None):
    """`kraken.loaders` entry point (`fn(path, tasks=None) -> list[BaseModel]`, kraken/models/loaders.py:28-43)."""
    if not _gpu_available():
        raise ValueError('kraken_b200: no CUDA device - leaving the file to the stock loaders')
    from kraken.models.loaders import load_coreml, load_safetensors
    errs = []
    for fn in (load_safetensors, load_coreml):
        try:
            models = fn(path, tasks=tasks)
            break
        except ValueError as e:
            errs.append(str(e))
    else:
        raise ValueError('; '.join(errs))
    from kraken.lib.vgsl import TorchVGSLModel as RefModel
    return [accelerate(m) if isinstance(m, RefModel) else m for m in models]


def make_registry_class():
    """The class behind the `kraken.models` name `TorchVGSLModelB200`: the reference class, accelerated when it is prepared for
    inference (weights are final by then; training keeps running on the stock PyTorch layers)."""
    from kraken.lib.vgsl import TorchVGSLModel as RefModel

    class TorchVGSLModelB200(RefModel):
        def prepare_for_inference(self, config):
            super().prepare_for_inference(config)
            dev = next(self.parameters()).device
            accelerate(self, device=str(dev) if dev.type == 'cuda' else 'cuda:0')

    return TorchVGSLModelB200


def __getattr__(name):            # `kraken_b200.accel:TorchVGSLModelB200` resolves lazily (needs kraken importable)
    if name == 'TorchVGSLModelB200':
        cls = make_registry_class()
        globals()[name] = cls
        return cls
    raise AttributeError(name)
