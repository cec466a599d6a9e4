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
