"""Host logic of the multi-GPU path with world_size 2 on the gloo backend (CPU): sharding, the single weight broadcast,
the single gather of decoded label blocks.  The per-rank engine is replaced by the CPU oracle here (tests only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kraken_b200.dist import broadcast_state_dict, gather_decoded, recognize_sharded, shard_batches

SPEC = '[1,16,0,1 Cr3,3,8 Mp2,2 S1(1x0)1,3 Lbx8 O1c12]'


def test_shard_batches_cover_everything_once():
    rng = np.random.default_rng(0)
    widths = rng.integers(20, 400, 103).tolist()
    for mode in ('arrival', 'bucketed'):
        for world in (1, 2, 8):
            sh = shard_batches(widths, world, 16, mode)
            flat = [i for r in sh for b in r for i in b]
            assert sorted(flat) == list(range(103))
            assert max(len(r) for r in sh) - min(len(r) for r in sh) <= 1
    arr = shard_batches(widths, 2, 16, 'arrival')
    assert arr[0][0] == list(range(16)) and arr[1][0] == list(range(16, 32))      # reference-identical padded batches
    buck = shard_batches(widths, 2, 16, 'bucketed')
    assert all(max(widths[i] for i in b) - min(widths[i] for i in b) <= 120 for r in buck for b in r)
    with pytest.raises(ValueError):
        shard_batches(widths, 2, 16, 'random')


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'oracle'))
    import vgsl_oracle as vo
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        om = vo.OracleModel(SPEC)
        w = om.init_like_reference(seed=0 if rank == 0 else 99)       # rank 1 starts with WRONG weights
        w = broadcast_state_dict({k: v.detach() for k, v in w.items()}, src=0)
        om.load(w)
        g = torch.Generator().manual_seed(3)
        widths = torch.randint(24, 90, (21,), generator=g).tolist()
        lines = [torch.rand(1, 16, wd, generator=g) for wd in widths]

        def rec(seqs, lens):
            return vo.rec_predict(om, seqs, lens)[3]
        out = recognize_sharded(rec, lines, batch_size=4, mode='arrival', stride=64)
        if rank == 0:
            q.put(out)
        # a rank that owns nothing still takes part in the gather
        res = gather_decoded([5] if rank == 1 else [], [[(3, 0, 1, 0.5)]] if rank == 1 else [], total=6, stride=4)
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_recognition_matches_single_process():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
    import vgsl_oracle as vo
    from kraken_b200.rpred import pad_batch
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=180)
    res = q.get(timeout=60)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process expectation with the rank-0 weights and the same arrival-order batches
    om = vo.OracleModel(SPEC)
    om.init_like_reference(seed=0)
    g = torch.Generator().manual_seed(3)
    widths = torch.randint(24, 90, (21,), generator=g).tolist()
    lines = [torch.rand(1, 16, wd, generator=g) for wd in widths]
    exp = []
    for i in range(0, 21, 4):
        seqs, lens = pad_batch(lines[i:i + 4])
        exp.extend(vo.rec_predict(om, seqs, lens)[3])
    assert len(out) == 21
    for a, b in zip(out, exp):
        assert [t[:3] for t in a] == [t[:3] for t in b]
        assert np.allclose([t[3] for t in a], [t[3] for t in b], atol=1e-6)
    assert res == [[], [], [], [], [], [(3, 0, 1, 0.5)]]


class _OracleRec:
    """Stands in for TorchSeqRecognizer.submit / collect on a box without a GPU: same block contract, oracle arithmetic."""

    def __init__(self, vo, om):
        self.vo, self.om, self._depth, self.q, self.n = vo, om, 0, {}, 0

    def set_pipeline_depth(self, d):
        self._depth = d

    def submit(self, line, lens=None, invert_max=None):
        self.n += 1
        t = self.n
        self.q[t] = self.vo.rec_predict(self.om, line, lens)[3]
        return t

    def collect(self, ticket, out=None):
        dec = self.q.pop(ticket)
        for k in ('labels', 'starts', 'ends', 'confs'):
            out[k][...] = 0
        for i, d in enumerate(dec):
            out['counts'][i] = len(d)
            for j, (l, s, e, c) in enumerate(d):
                out['labels'][i, j], out['starts'][i, j], out['ends'][i, j], out['confs'][i, j] = l, s, e, c
        return out


def _worker_blocks(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'oracle'))
    import vgsl_oracle as vo
    from kraken_b200.dist import recognize_sharded_blocks
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        om = vo.OracleModel(SPEC)
        om.init_like_reference(seed=0)
        g = torch.Generator().manual_seed(11 + rank)                       # every rank owns DIFFERENT lines (its shard)
        mine = [(torch.rand(4, 1, 16, 60, generator=g), torch.tensor([60, 33, 60, 17])) for _ in range(3)]
        blocks, gathered = recognize_sharded_blocks(_OracleRec(vo, om), mine, stride=15, depth=2, device=None, dst=0)
        if rank == 0:
            q.put([t.numpy().copy() for t in gathered])
        else:
            assert gathered is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_result_blocks_single_gather():
    """The throughput path of the multi-GPU job: result blocks written in place, ONE gather of the whole run (dist.ResultBlocks)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
    import vgsl_oracle as vo
    from kraken_b200.dist import ResultBlocks
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_blocks, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(got) == 2 and got[0].shape == (3, 4 * (1 + 4 * 15))
    om = vo.OracleModel(SPEC)
    om.init_like_reference(seed=0)
    for rank in range(2):
        g = torch.Generator().manual_seed(11 + rank)
        rb = ResultBlocks(3, 4, 15, pin=False)
        rb.a[...] = got[rank]
        for i in range(3):
            x, lens = torch.rand(4, 1, 16, 60, generator=g), torch.tensor([60, 33, 60, 17])
            exp = vo.rec_predict(om, x, lens)[3]
            v = rb.views(i)
            for j, d in enumerate(exp):
                assert int(v['counts'][j]) == len(d)
                assert [(int(v['labels'][j, k]), int(v['starts'][j, k]), int(v['ends'][j, k])) for k in range(len(d))] == [t[:3] for t in d]
                assert np.allclose(v['confs'][j, :len(d)], [t[3] for t in d], atol=1e-6)
