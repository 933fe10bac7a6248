"""
Multi-GPU recognition: lines are independent (kraken/lib/vgsl/rpred.py:129-131 only pads and stacks), so the path
shards with NO data-path collective.  One process per GPU (torch.distributed, NCCL over NVLink):

    1. `broadcast_state_dict`   - ONE broadcast of the packed fp32 weight blob from rank 0 at load time
    2. `shard_batches`          - deterministic partition of the line list into batches, dealt to ranks
    3. every rank runs its batches through its own engine replica (`kb_recognize`)
    4. `gather_decoded`         - ONE gather of the fixed-stride label blocks to rank 0 at the end

The reference has no inference data parallelism at all; this module is the new capability of 8e.
Works with the `gloo` backend on CPU tensors too, which is how the host logic is tested without GPUs.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

__all__ = ['shard_batches', 'broadcast_state_dict', 'gather_decoded', 'recognize_sharded', 'bind_to_gpu_numa', 'ResultBlocks',
           'recognize_sharded_blocks']


def shard_batches(widths: Sequence[int], world_size: int, batch_size: int, mode: str = 'arrival') -> list[list[list[int]]]:
    """Returns, per rank, a list of batches (lists of line indices).

    mode 'arrival'  - batches are consecutive runs of `batch_size` lines in arrival order, exactly the padded batches the
                      reference forms (rpred.py:115-121), dealt round-robin to ranks.  Results are then identical to a
                      single-GPU / reference run (padding semantics are per batch).
    mode 'bucketed' - lines are sorted by width first (minimises padding), batches dealt round-robin so every rank sees
                      the same width mix.  Faster; padded-batch composition differs from the reference's.
    """
    n = len(widths)
    order = list(range(n)) if mode == 'arrival' else sorted(range(n), key=lambda i: (widths[i], i))
    if mode not in ('arrival', 'bucketed'):
        raise ValueError(f'unknown sharding mode {mode}')
    batches = [order[i:i + batch_size] for i in range(0, n, batch_size)]
    if mode == 'bucketed':
        # widest batches first so the expensive work is dealt before the cheap tail
        batches.sort(key=lambda b: -max(widths[i] for i in b))
    out: list[list[list[int]]] = [[] for _ in range(world_size)]
    for k, b in enumerate(batches):
        out[k % world_size].append(b)
    return out


def broadcast_state_dict(state_dict: dict, src: int = 0, device: Optional[torch.device] = None) -> dict:
    """One collective for all weights: flatten -> broadcast -> unflatten (keys/shapes are known on every rank
    because every rank parsed the same VGSL spec)."""
    keys = list(state_dict.keys())
    flat = torch.cat([torch.as_tensor(state_dict[k]).float().flatten() for k in keys])
    if device is not None:
        flat = flat.to(device)
    dist.broadcast(flat, src=src)
    flat = flat.cpu()
    out, off = {}, 0
    for k in keys:
        t = torch.as_tensor(state_dict[k])
        out[k] = flat[off:off + t.numel()].view(t.shape).clone()
        off += t.numel()
    return out


def _pack(indices: Sequence[int], decoded: Sequence[Sequence[tuple]], stride: int) -> torch.Tensor:
    """[(label, start, end, conf)]* per line -> int32 [n, 2 + 4*stride]: line index, count, then 4 planes."""
    n = len(indices)
    t = np.zeros((n, 2 + 4 * stride), np.int32)
    for r, (idx, d) in enumerate(zip(indices, decoded)):
        k = len(d)
        if k > stride:
            raise ValueError('decoded sequence longer than the pack stride')
        t[r, 0], t[r, 1] = idx, k
        if k:
            a = np.asarray([(x[0], x[1], x[2]) for x in d], np.int32)
            t[r, 2:2 + k] = a[:, 0]
            t[r, 2 + stride:2 + stride + k] = a[:, 1]
            t[r, 2 + 2 * stride:2 + 2 * stride + k] = a[:, 2]
            t[r, 2 + 3 * stride:2 + 3 * stride + k] = np.asarray([x[3] for x in d], np.float32).view(np.int32)
    return torch.from_numpy(t)


def _unpack(t: torch.Tensor, stride: int):
    a = t.cpu().numpy()
    out = {}
    for row in a:
        idx, k = int(row[0]), int(row[1])
        if idx < 0:
            continue
        conf = row[2 + 3 * stride:2 + 3 * stride + k].view(np.float32)
        out[idx] = [(int(row[2 + j]), int(row[2 + stride + j]), int(row[2 + 2 * stride + j]), float(conf[j])) for j in range(k)]
    return out


def gather_decoded(indices: Sequence[int], decoded: Sequence[Sequence[tuple]], total: int, stride: int, dst: int = 0,
                   device: Optional[torch.device] = None):
    """ONE gather of all ranks' label blocks to `dst`; returns the full in-order list there, None elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    per_rank = (total + world - 1) // world + 1
    # ranks may own different numbers of lines; pad to a common row count with index -1
    rows = max(per_rank, len(indices))
    cnt = torch.tensor([rows], dtype=torch.int64, device=device) if device is not None else torch.tensor([rows], dtype=torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
    rows = int(cnt.item())
    block = torch.full((rows, 2 + 4 * stride), -1, dtype=torch.int32)
    if len(indices):
        block[:len(indices)] = _pack(indices, decoded, stride)
    if device is not None:
        block = block.to(device)
    bufs = [torch.empty_like(block) for _ in range(world)] if rank == dst else None
    dist.gather(block, bufs, dst=dst)
    if rank != dst:
        return None
    merged = {}
    for b in bufs:
        merged.update(_unpack(b, stride))
    return [merged.get(i, []) for i in range(total)]


def recognize_sharded(recognize_batch: Callable, lines: Sequence[torch.Tensor], batch_size: int = 64, mode: str = 'arrival',
                      stride: Optional[int] = None, device: Optional[torch.device] = None, dst: int = 0):
    """`recognize_batch(padded[N,C,H,W], lens) -> list[list[(label,start,end,conf)]]` is this rank's engine
    (e.g. TorchSeqRecognizer.predict_labels).  Every rank passes the SAME `lines`; rank `dst` gets all results."""
    from .rpred import pad_batch
    world, rank = dist.get_world_size(), dist.get_rank()
    widths = [int(l.shape[2]) for l in lines]
    mine = shard_batches(widths, world, batch_size, mode)[rank]
    idxs, decs = [], []
    for b in mine:
        seqs, lens = pad_batch([lines[i] for i in b])
        out = recognize_batch(seqs, lens)
        idxs.extend(b)
        decs.extend(out)
    if stride is None:
        stride = max(1, max(widths) if widths else 1)
    return gather_decoded(idxs, decs, len(lines), stride, dst=dst, device=device)


# ---- throughput path: result blocks instead of Python tuples ---------------------------------------------------------------------
def bind_to_gpu_numa(device_index: int) -> Optional[list]:
    """Restricts this process to the CPU cores that are local to GPU `device_index` (NVML's ideal CPU affinity, intersected with
    the cores the process may use) - call it before pinned buffers are allocated and before worker threads start, so that the
    staging memory is first-touched on the GPU's NUMA node and the H2D copies do not cross the socket interconnect.  Returns the
    core list, or None when NVML / affinity control is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        cvd = [x for x in os.environ.get('CUDA_VISIBLE_DEVICES', '').split(',') if x.strip() != '']
        idx = int(cvd[device_index]) if cvd and cvd[device_index].isdigit() else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        ideal = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cores = sorted(ideal & allowed)
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:
        return None


class ResultBlocks:
    """The decoded label blocks of a run of `nbatches` batches of `batch` lines in ONE pinned int32 buffer
    [batch index][counts (B) | labels (B x stride) | starts | ends | confs (float32 bits)]: the engine writes every batch straight
    into its slice (`views(i)` are the arrays `TorchSeqRecognizer.collect(out=...)` / `_recognize_raw(out=...)` take) and the run's
    single gather ships the buffer as it is."""

    def __init__(self, nbatches: int, batch: int, stride: int, pin: bool = True):
        self.nbatches, self.batch, self.stride = nbatches, batch, stride
        self.words = batch * (1 + 4 * stride)
        self.t = torch.empty((max(nbatches, 1), self.words), dtype=torch.int32)
        if pin and torch.cuda.is_available():
            self.t = self.t.pin_memory()
        self.a = self.t.numpy()

    def views(self, i: int) -> dict:
        b, o, bt = self.a[i], self.batch, self.batch * self.stride
        shp = (self.batch, self.stride)
        return {'counts': b[:o], 'labels': b[o:o + bt].reshape(shp), 'starts': b[o + bt:o + 2 * bt].reshape(shp),
                'ends': b[o + 2 * bt:o + 3 * bt].reshape(shp), 'confs': b[o + 3 * bt:o + 4 * bt].view(np.float32).reshape(shp)}

    def gather(self, dst: int = 0, device: Optional[torch.device] = None):
        """ONE collective for the whole run (NCCL gather over NVLink when `device` is a CUDA device; gloo on CPU tensors).  Returns the
        list of all ranks' buffers on `dst`, None elsewhere.  Every rank must hold the same number of batches."""
        world, rank = dist.get_world_size(), dist.get_rank()
        t = self.t.to(device, non_blocking=True) if device is not None else self.t
        out = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, out, dst=dst)
        return out


def recognize_sharded_blocks(rec, my_batches: Sequence, stride: int, depth: int = 4, device: Optional[torch.device] = None, dst: int = 0,
                             blocks: Optional[ResultBlocks] = None):
    """This rank's share of a job: `my_batches` = [(lines[B,C,H,W], lens)] (host pinned or device tensors, all with the same B) go
    through the engine's asynchronous pipeline (`depth` batches in flight from this one thread), results land in a `ResultBlocks`
    buffer, and ONE gather delivers every rank's buffer to `dst`.  No collective on the data path.  Returns (blocks, gathered)."""
    nb = len(my_batches)
    if blocks is None:
        blocks = ResultBlocks(nb, int(my_batches[0][0].shape[0]) if nb else 1, stride)
    if getattr(rec, '_depth', 0) != depth:
        rec.set_pipeline_depth(depth)
    pend = []
    for i, b in enumerate(my_batches):
        if len(pend) == depth:
            j, t = pend.pop(0)
            rec.collect(t, out=blocks.views(j))
        pend.append((i, rec.submit(*b)))
    while pend:
        j, t = pend.pop(0)
        rec.collect(t, out=blocks.views(j))
    gathered = blocks.gather(dst, device) if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 else [blocks.t]
    return blocks, gathered
