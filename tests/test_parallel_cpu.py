"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: sample sharding + the one all-reduce of the
(sum NLL, count) payload (nfb200/parallel.py).  The per-shard numbers come from the CPU oracle -- the CUDA path is
covered by the -m gpu tests; here only the partition / reduction arithmetic is under test."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_partition_every_row_exactly_once():
    from nfb200 import parallel
    for n in (0, 1, 7, 256, 2048, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            bounds = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            for (a0, a1), (b0, b1) in zip(bounds, bounds[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in bounds]
            assert max(sizes) - min(sizes) <= 1  # balanced


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from nfb200 import parallel
        from oracle import flow_oracle as O
        from tests import _golden
        meta, a, sd = _golden.load('model_realnvp_64d')
        spec = O.stack_spec('realnvp', tuple(meta['dims']), meta['datatype'], meta['layers'])
        x = a['x'][:61]  # 61 rows: uneven split 31 + 30
        xs = parallel.shard_rows(x, rank, world)
        with torch.no_grad():
            z, ldj = O.stack_forward(spec, sd, xs)
        rows = O.nll_rows(z, ldj)
        total = torch.tensor([float(rows.sum()), float(rows.numel())], dtype=torch.float64)
        bpd = parallel.global_bits_per_dim(total, 64)
        if rank == 0:
            with torch.no_grad():
                zf, lf = O.stack_forward(spec, sd, x)
            torch.save({'bpd_sharded': bpd, 'bpd_full': O.bits_per_dim(zf, lf), 'count': float(total[1])}, out_path)
    finally:
        dist.destroy_process_group()


def test_sharded_nll_allreduce_matches_single_process(tmp_path):
    out = str(tmp_path / 'res.pt')
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res['count'] == 31.0  # rank 0's own shard stays untouched (all-reduce works on a clone)
    assert abs(res['bpd_sharded'] - res['bpd_full']) <= 1e-12 * abs(res['bpd_full'])


def test_allreduce_is_noop_without_process_group():
    from nfb200 import parallel
    t = torch.tensor([3.0, 2.0], dtype=torch.float64)
    assert parallel.allreduce_nll(t) is t
    assert parallel.global_bits_per_dim(t, 1) == pytest.approx(1.5 / 0.6931471805599453)


# ---- data-parallel training step: flat-bucket gradient all-reduce (nfb200.parallel.train_step) -------------------------
class _OracleFlow(torch.nn.Module):
    """A CPU stand-in with the product models' interface (``nll(x) -> (rows, total)``), built on the oracle stack, so the
    host-side logic of train_step / allreduce_gradients can run under gloo.  (The CUDA layers themselves: -m gpu tests.)"""

    def __init__(self, spec, sd):
        super().__init__()
        self.spec = spec
        self.keys = [k for k, v in sd.items() if v.is_floating_point() and k.split('.')[-1] in ('s_log_scale', 's_bias', 'weight_g', 'weight_v', 'bias', 'weight')]
        self.params = torch.nn.ParameterList([torch.nn.Parameter(sd[k].clone()) for k in self.keys])
        self.rest = {k: v for k, v in sd.items() if k not in self.keys}

    def nll(self, x):
        from oracle import flow_oracle as O
        sd = dict(self.rest)
        sd.update({k: p for k, p in zip(self.keys, self.params)})
        z, ldj = O.stack_forward(self.spec, sd, x)
        zf = z.reshape(z.shape[0], -1)
        rows = 0.5 * (zf * zf).sum(1) + 0.5 * zf.shape[1] * 1.8378770664093453 - ldj
        total = torch.tensor([float(rows.sum()), float(rows.numel())], dtype=torch.float64)
        return rows, total


def _train_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from nfb200 import parallel
        from oracle import flow_oracle as O
        from tests import _golden
        meta, a, sd = _golden.load('model_realnvp_64d')
        spec = O.stack_spec('realnvp', tuple(meta['dims']), meta['datatype'], meta['layers'])
        x = a['x'][:37]  # uneven split 19 + 18
        net = _OracleFlow(spec, sd)
        opt = torch.optim.SGD(net.parameters(), lr=1e-3)
        losses = [parallel.train_step(net, opt, parallel.shard_rows(x, rank, world)) for _ in range(3)]
        if rank == 0:
            torch.save({'losses': losses, 'params': [p.detach().clone() for p in net.params]}, out_path)
    finally:
        dist.destroy_process_group()


def test_sharded_train_step_matches_single_process(tmp_path):
    """Two ranks with uneven shards take the same optimisation steps as one process on the whole batch."""
    sys.path.insert(0, ROOT)
    from nfb200 import parallel
    from oracle import flow_oracle as O
    from tests import _golden
    out = str(tmp_path / 'train.pt')
    mp.spawn(_train_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    meta, a, sd = _golden.load('model_realnvp_64d')
    spec = O.stack_spec('realnvp', tuple(meta['dims']), meta['datatype'], meta['layers'])
    net = _OracleFlow(spec, sd)
    opt = torch.optim.SGD(net.parameters(), lr=1e-3)
    torch.set_num_threads(2)
    losses = [parallel.train_step(net, opt, a['x'][:37]) for _ in range(3)]
    for l0, l1 in zip(losses, res['losses']):
        assert abs(l0 - l1) <= 1e-5 * abs(l0)
    assert losses[-1] < losses[0]
    for p0, p1 in zip(net.params, res['params']):
        assert torch.allclose(p0, p1, rtol=1e-4, atol=1e-6)
