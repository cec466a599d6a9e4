"""Sample-sharded evaluation across the GPUs of one box.

Every layer is per-sample independent in eval mode, so rank r simply owns rows [r*B/N, (r+1)*B/N) of the global
batch with replicated weights; there is NO data-path collective.  The only exchange is one all-reduce (SUM) of the
2-element fp64 payload (sum of per-sample NLL, sample count) -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist

from . import _lib as L
from .likelihood import LN2


def shard_bounds(n_rows, rank, world_size):
    """Contiguous, balanced row range of `rank`: the first n_rows % world_size ranks get one extra row."""
    base, extra = divmod(n_rows, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(x, rank, world_size):
    lo, hi = shard_bounds(x.size(0), rank, world_size)
    return x[lo:hi]


def allreduce_nll(total, group=None):
    """In-place SUM all-reduce of the (sum NLL, count) payload; no-op without an initialised process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return total


def global_bits_per_dim(total, D, group=None):
    total = allreduce_nll(total.clone(), group)
    s, n = (float(v) for v in total.tolist())
    return s / n / (D * LN2)


def sharded_bits_per_dim(model, x_local, group=None):
    """bits/dim of the GLOBAL batch given this rank's shard `x_local` (already on this rank's GPU)."""
    _, total = model.nll(x_local)
    return global_bits_per_dim(total, x_local[0].numel(), group)


# ---- cross-sample statistics on a sharded batch (the places where samples are NOT independent) ----------


def channel_moments(z):
    """device double[2C + 1] = (sum x per channel, sum x^2 per channel, element count per channel) of this rank's shard."""
    from . import _lib as L
    z = L.dev(z, 'z')
    B, C = z.size(0), z.size(1)
    HW = z[0, 0].numel() if z.dim() > 2 else 1
    out = torch.empty(2 * C + 1, device=z.device, dtype=torch.float64)
    L.check(L.lib().nfb_channel_moments(L.ptr(z), out.data_ptr(), B, C, HW, L.stream()))
    out[2 * C] = float(B * HW)
    return out


def finalize_moments(m):
    """(mean, biased variance, unbiased variance, n) per channel from (all-reduced) moments, in fp64."""
    C = (m.numel() - 1) // 2
    n = m[2 * C]
    mean = m[:C] / n
    ss = torch.clamp(m[C:2 * C] - n * mean * mean, min=0.0)
    return mean, ss / n, ss / torch.clamp(n - 1, min=1.0), n


def actnorm_init_sharded(layer, z_local, group=None):
    """ActNorm's data-dependent init (modules.py:238-244) from the GLOBAL batch when every rank holds a shard: one
    all-reduce of 2C+1 doubles; every rank ends up with identical log_scale / bias."""
    m = channel_moments(z_local)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(m, op=dist.ReduceOp.SUM, group=group)
    mean, _, var_unbiased, _ = finalize_moments(m)
    with torch.no_grad():
        layer.log_scale.data.copy_(torch.log(torch.sqrt(var_unbiased) + layer.eps).float().view(layer.dimensions))
        layer.bias.data.copy_(mean.float().view(layer.dimensions))
    layer.initialized = True
    return layer


def batchnorm_stats_sharded(layer, x_local, group=None):
    """Train-mode statistics of the flow BatchNorm (modules.py:285-294) from the global batch: batch_mean / batch_var
    (+eps) and the running-statistics update, identical on every rank."""
    m = channel_moments(x_local)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(m, op=dist.ReduceOp.SUM, group=group)
    mean, var_biased, _, _ = finalize_moments(m)
    with torch.no_grad():
        layer.batch_mean.copy_(mean.float().view(layer.dimensions))
        layer.batch_var.copy_((var_biased.float() + layer.eps).view(layer.dimensions))
        layer.running_mean.mul_(1.0 - layer.momentum).add_(layer.batch_mean * layer.momentum)
        layer.running_var.mul_(1.0 - layer.momentum).add_(layer.batch_var * layer.momentum)
    return layer


# ---- data-parallel training: one process per GPU, replicated weights, sample-sharded batch ----------


def allreduce_gradients(model, group=None, world_size=None):
    """SUM all-reduce of every parameter gradient as ONE flat bucket (NCCL over NVLink / NVSwitch: a single launch-
    latency-bound collective instead of one per tensor).  No-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    if (world_size or dist.get_world_size(group)) <= 1:
        return
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def train_step(model, optimizer, x_local, group=None):
    """One optimisation step of main.py:78-92 on this rank's shard of the global batch.

    loss = mean over the GLOBAL batch of -(log N(z; 0, I) + ldj): every rank back-propagates sum(local NLL) / B_global
    and the gradients are summed across ranks.  Cross-sample state: ActNorm's first-call init and the flow BatchNorm's
    train-mode statistics use the all-reduced moments of the global batch (flows.modules.SYNC_STATS), so the replicas
    hold identical bijection parameters and buffers; the nn.BatchNorm layers INSIDE the conditioners keep per-rank batch
    statistics, as in plain DDP -- with them the update equals the single-process step up to that difference (exactly for
    stacks whose conditioners run in eval mode).
    Returns the global mean NLL as a float."""
    rows, total = model.nll(x_local)
    total = allreduce_nll(total.clone(), group)
    loss = rows.sum() / total[1].to(rows.dtype)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    allreduce_gradients(model, group)
    optimizer.step()
    s, n = (float(v) for v in total.tolist())
    return s / n


class GraphedTrainStep:
    """The training step as CUDA-graph replays: graph 1 = train-mode forward + loss + backward (every libnfb200 kernel,
    any library call of a conditioner that is not covered natively, and torch's glue captured once), then the gradient
    all-reduce on the flat bucket, then a fused multi-tensor optimizer step (``capturable`` Adam inside graph 2).  The eager step of a Glow K=32
    is dominated by Python / launch overhead (thousands of small launches); replaying removes it.

    All parameter gradients are views into ONE flat fp32 buffer, so the all-reduce needs no gather / scatter copies and
    zeroing the gradients is a single memset node of the graph.

    Usage::

        step = GraphedTrainStep(model, torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True), example_batch)
        loss = step(x_local)          # device tensor holding the global mean NLL; no host sync
    """

    def __init__(self, model, optimizer, example, group=None, warmup=2):
        if not example.is_cuda:
            raise RuntimeError('nfb200: GraphedTrainStep needs CUDA tensors (no CPU fallback)')
        self.model, self.optimizer, self.group = model, optimizer, group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        params = [p for p in model.parameters() if p.requires_grad]
        self.flat = torch.zeros(sum(p.numel() for p in params), device=example.device, dtype=torch.float32)
        off = 0
        for p in params:  # gradients live inside the bucket for good
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.x = example.clone()
        self.total = torch.zeros(2, device=example.device, dtype=torch.float64)   # (sum NLL, count), all-reduced
        self.count = torch.ones((), device=example.device, dtype=torch.float32)   # global batch size, device-resident
        self.loss = torch.zeros((), device=example.device, dtype=torch.float32)
        self._sync_count()
        was_training = model.training
        with torch.no_grad():
            model.eval()   # eval: no BatchNorm running-statistic update from this forward
            model(self.x)  # ActNorm's data-dependent init (modules.py:238-244) happens here if it has not yet
            model.train(was_training)
        # warm-up steps (allocator, cuDNN plans, optimizer state) must not move the model: snapshot, restore below
        snap = {k: v.clone() for k, v in model.state_dict().items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):  # outside capture; at least once so the optimizer state exists
                self._fwd_bwd()
                self._reduce()
                optimizer.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            for k, v in model.state_dict().items():
                v.copy_(snap[k])  # in place: the captured graphs keep pointing at the same storage
            for st in optimizer.state.values():  # fresh optimizer state (step = 0, moments = 0), same storage
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
        self.g_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fb):
            self._fwd_bwd()
        self.g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_opt):
            optimizer.step()

    def _sync_count(self):
        n = torch.tensor([float(self.x.size(0))], device=self.x.device, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(n, group=self.group)
        self.count.copy_(n[0].float())

    def _fwd_bwd(self):
        self.flat.zero_()
        rows, total = self.model.nll(self.x)
        self.total.copy_(total)
        (rows.sum() / self.count).backward()  # gradients accumulate into the views of the flat bucket

    def _reduce(self):
        if self.world > 1:
            dist.all_reduce(self.flat, group=self.group)
            dist.all_reduce(self.total, group=self.group)
        self.loss.copy_((self.total[0] / self.total[1]).float())

    def __call__(self, x_local):
        if x_local.shape != self.x.shape:
            raise RuntimeError('nfb200: GraphedTrainStep was captured for batches of shape %s' % (tuple(self.x.shape), ))
        self.x.copy_(x_local, non_blocking=True)
        self.g_fb.replay()
        self._reduce()
        self.g_opt.replay()
        # the replayed optimizer step rewrote the parameters without bumping their _version counters: every derived-weight
        # cache (folded conditioner weights, W / W^-1, WeightNorm folds) must be rebuilt by the next eval forward
        L.bump_weights_epoch()
        return self.loss
