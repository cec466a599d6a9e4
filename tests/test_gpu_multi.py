"""Two-GPU tests (NCCL, one process per GPU, spawned from the test): run with ``-m gpu`` on a box with >= 2 devices;
skipped otherwise.  SURVEY.md 4 item 5 / 8e: a sharded global batch gives the single-GPU bits/dim; cross-sample
statistics (ActNorm init) equal the full-batch ones; replicas stay identical after a training step."""
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


def _need2():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 CUDA devices')


def _worker(rank, world, port, what, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import nfb200
        from nfb200 import parallel
        dev = torch.device('cuda', rank)
        cfg = types.SimpleNamespace(layers=2, mixtures=4)
        dims = (3, 16, 16)
        g = torch.Generator().manual_seed(0)
        x_all = torch.rand(38, *dims, generator=g)  # 19 / 19 split
        torch.manual_seed(0)
        net = nfb200.Glow(dims, 'image', cfg).to(dev)
        if what == 'bpd':
            # eval-mode density of the global batch: initialise every replica identically from the FULL batch, then shard
            net.eval()
            with torch.no_grad():
                nfb200.flows.modules.SYNC_STATS = False
                net(x_all.to(dev))  # every rank: same data -> same ActNorm init
                full = net.bits_per_dim(x_all.to(dev))
                sharded = parallel.sharded_bits_per_dim(net, parallel.shard_rows(x_all, rank, world).to(dev))
            out[rank] = (full, sharded)
        elif what == 'actnorm':
            # ActNorm init from shards (all-reduced moments) == init from the full batch
            net.eval()
            with torch.no_grad():
                net(parallel.shard_rows(x_all, rank, world).to(dev))  # SYNC_STATS default: global statistics
                a = {k: v.detach().cpu().clone() for k, v in net.state_dict().items() if k.endswith(('log_scale', 'bias'))}
                torch.manual_seed(0)
                ref = nfb200.Glow(dims, 'image', cfg).to(dev).eval()
                nfb200.flows.modules.SYNC_STATS = False
                ref(x_all.to(dev))
                b = {k: v.detach().cpu().clone() for k, v in ref.state_dict().items() if k.endswith(('log_scale', 'bias'))}
            out[rank] = max(float((a[k] - b[k]).abs().max()) for k in a)
        elif what == 'train':
            net.train()
            opt = torch.optim.Adam(net.parameters(), lr=1e-3)
            for _ in range(2):
                parallel.train_step(net, opt, parallel.shard_rows(x_all, rank, world).to(dev))
            flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()] +
                             [b.detach().reshape(-1).float() for k, b in net.named_buffers()
                              if k.endswith(('log_scale', 'bias')) or 'running' not in k and 'num_batches' not in k])
            other = flat.clone()
            dist.broadcast(other, src=0)
            out[rank] = float((flat - other).abs().max())
    finally:
        dist.destroy_process_group()


def _run(what):
    import torch.multiprocessing as mp
    _need2()
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29650 + (os.getpid() % 200)
    mp.spawn(_worker, args=(2, port, what, out), nprocs=2, join=True)
    return dict(out)


def test_sharded_bits_per_dim_equals_single_gpu():
    out = _run('bpd')
    for rank in (0, 1):
        full, sharded = out[rank]
        assert abs(full - sharded) <= 1e-12 * abs(full), (full, sharded)
    assert out[0][1] == out[1][1]  # both ranks hold the same all-reduced value


def test_actnorm_init_from_shards_equals_full_batch():
    out = _run('actnorm')
    assert max(out.values()) < 2e-6, out


def test_replicas_identical_after_training_steps():
    out = _run('train')
    assert out[1] == 0.0, out  # rank 1's bijection parameters and ActNorm state are bit-identical to rank 0's
