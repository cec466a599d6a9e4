#!/usr/bin/env python
"""bench.py -- samples/sec of the flow forward + log-det + NLL hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl nfb200|reference] [--workload NAME]

One "step" = one eval-mode pass of the whole stack over one batch of synthetic inputs:
``model.forward(x) -> (z, log_df_dz)`` followed by the Gaussian NLL reduction (and, for N > 1, the all-reduce of
the 2-element (sum NLL, count) payload -- the only collective).  Default workload: BASELINE.json configs[1],
Glow K=32 L=3 on 32x32x3, batch 256 per GPU (weak scaling: every rank owns its own 256 samples).

Prints ONE JSON line (rank 0).  `value` = whole-job samples/s with inputs resident in HBM (CUDA-graph replay,
CUDA-event timing, max over ranks); `e2e` = the same through the public nn.Module API with pinned HOST buffers
(H2D copy of x and D2H read of the NLL inside the timed region); `roofline` = dominant hand-written kernel, timed
live with CUDA events; `cpu_baseline` = the CPU oracle (port of the reference's PyTorch CPU path) on the box's host
cores.  `--impl reference` times that CPU path alone.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

L2_BYTES = 126 * 1024 * 1024

WORKLOADS = {
    # name: (model, dims, datatype, cfg kwargs, batch per GPU, description)
    'glow32': ('glow', (3, 32, 32), 'image', dict(layers=32, mixtures=4), 256,
               'Glow K=32 L=3 on 32x32x3 synthetic images, batch 256 per GPU (BASELINE.json configs[1])'),
    'flowpp32': ('flowpp', (3, 32, 32), 'image', dict(layers=32, mixtures=8), 256,
                 'Flow++ logistic-mixture coupling on 32x32x3 synthetic, batch 256 (configs[2])'),
    'realnvp64_rqs': ('realnvp', (64, ), None, dict(layers=8, mixtures=4, coupling='rqs'), 65536,
                      'RealNVP 8 RQ-spline couplings on 64-dim synthetic tabular, batch 65536 (configs[3])'),
    'realnvp64': ('realnvp', (64, ), None, dict(layers=8, mixtures=4), 65536,
                  'RealNVP 8 affine couplings on 64-dim synthetic tabular, batch 65536 (configs[3], affine proxy)'),
    'realnvp2': ('realnvp', (2, ), None, dict(layers=6, mixtures=4), 512,
                 'RealNVP 6 affine couplings on 2D, batch 512 (configs[0])'),
    'glow64': ('glow', (3, 64, 64), 'image', dict(layers=48, mixtures=4), 256,
               'Glow K=48 L=4 on 64x64x3 synthetic, 256 per GPU = 2048 over 8 GPUs (configs[4])'),
}


def make_inputs(dims, datatype, batch, seed):
    g = torch.Generator().manual_seed(seed)
    if datatype == 'image':
        return torch.rand((batch, ) + tuple(dims), generator=g)
    return torch.randn((batch, ) + tuple(dims), generator=g)


def oracle_spec(wl):
    from oracle import flow_oracle as O
    model, dims, datatype, cfg, _, _ = WORKLOADS[wl]
    return O.stack_spec(model, dims, datatype, cfg['layers'], cfg.get('mixtures', 4), cfg.get('coupling'))


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': max(mx), 'power_w_max': max(power), 'samples': len(sm),
                'reasons': sorted(reasons)}


def cpu_forward_timer(wl, sd, x, threads):
    """Time the CPU oracle (port of the reference's PyTorch CPU path) on a bounded sample."""
    from oracle import flow_oracle as O
    spec = oracle_spec(wl)
    torch.set_num_threads(threads)
    with torch.no_grad():
        t0 = time.perf_counter()
        z, ldj = O.stack_forward(spec, sd, x)  # warm-up
        warm = time.perf_counter() - t0
        best = warm
        reps = 0
        budget = time.perf_counter() + 20.0
        while reps < 3 and time.perf_counter() < budget:
            t0 = time.perf_counter()
            z, ldj = O.stack_forward(spec, sd, x)
            best = min(best, time.perf_counter() - t0)
            reps += 1
    return best, O.bits_per_dim(z, ldj), reps


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the Python reference
    cannot travel to the GPU box), all host threads, same workload/metric."""
    if rank != 0:
        return
    import nfb200  # only for a state dict of the right architecture (random init, CPU tensors; no CUDA calls)
    from oracle import flow_oracle as O
    model, dims, datatype, cfg, batch, desc = WORKLOADS[args.workload]
    torch.manual_seed(0)
    net = getattr(nfb200, {'glow': 'Glow', 'flowpp': 'Flowpp', 'realnvp': 'RealNVP'}[model])(
        dims, datatype, types.SimpleNamespace(**cfg))
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    spec = oracle_spec(args.workload)
    # bounded sample: size each step so the whole run stays within ~2 minutes
    probe_n = min(batch, 16)
    xp = make_inputs(dims, datatype, probe_n, 123)
    with torch.no_grad():
        O.stack_forward(spec, sd, xp)
        t0 = time.perf_counter()
        O.stack_forward(spec, sd, xp)
        per_sample = (time.perf_counter() - t0) / probe_n
    n_steps = args.steps + args.warmup
    sample = int(max(1, min(batch, 120.0 / max(per_sample * n_steps, 1e-9))))
    x = make_inputs(dims, datatype, sample, 0)
    with torch.no_grad():
        for _ in range(args.warmup):
            O.stack_forward(spec, sd, x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            z, ldj = O.stack_forward(spec, sd, x)
        dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    out = {
        'impl': 'reference', 'metric': 'samples/sec (fwd+logdet)', 'value': value, 'unit': 'samples/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': desc, 'batch_per_step': sample},
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
                         'sample': '%d samples per step of the %d-sample batch, torch %s CPU' %
                                   (sample, batch, torch.__version__)},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(out)


def time_kernel_stream(fn, iters, flush=None):
    """Average device time of fn() in ms, CUDA events on the current stream, optional L2 flush between calls."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        if flush is not None:
            flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return sum(ts) / len(ts), ts[len(ts) // 2], ts[0]


def _ncu_traffic(key):
    """dram__bytes_read+write per launch from the committed ncu --set full summary (profiles/ncu_summary.json)."""
    p = os.path.join(ROOT, 'profiles', 'ncu_summary.json')
    try:
        with open(p) as f:
            return json.load(f).get(key, {}).get('dram_bytes_per_launch')
    except (OSError, ValueError):
        return None


def coupling_roofline(peaks):
    """Streaming-size run of the fused affine coupling + log-det kernel (the kernel BASELINE.json's target names):
    z, params and out are 201 MB each (604 MB > 126 MB L2), L2 flushed between launches, CUDA events on the
    launching stream.  Headline = checkerboard split at the cfg-2 first-level shape; the other splits ride along."""
    import nfb200._lib as L
    B, D = 16384, 3072
    z = torch.randn(B, D, device='cuda')
    params = torch.randn(B, D, device='cuda')
    out = torch.empty_like(z)
    ldj = torch.zeros(B, device='cuda')
    a = torch.tensor([0.3], device='cuda')
    b = torch.tensor([0.01], device='cuda')
    flush = torch.empty(L2_BYTES * 2 // 4, device='cuda', dtype=torch.float32)
    st = L.stream()
    alg = (12 * D + 8) * B  # z in 4D + (t, s) 4D + z out 4D + log-det read-modify-write
    peak = peaks['hbm_gbs']
    res = {}
    for name, (C, H, W, mode) in {'checkerboard 3x32x32': (3, 32, 32, L.SPLIT_CHECKER),
                                  'channelwise 12x16x16': (12, 16, 16, L.SPLIT_CHANNEL),
                                  '1-D 3072': (3072, 1, 1, L.SPLIT_1D)}.items():
        def run():
            L.check(L.lib().nfb_affine_coupling_fwd(z.data_ptr(), out.data_ptr(), params.data_ptr(), ldj.data_ptr(),
                                                    ldj.data_ptr(), a.data_ptr(), b.data_ptr(), B, C, H, W, mode, 0, st))
        for _ in range(3):
            run()
        mean, med, best = time_kernel_stream(run, 20, flush)
        res[name] = {'ms_per_launch': med, 'achieved': alg / (med * 1e-3) / 1e9}
    head = res['checkerboard 3x32x32']
    return {'kernel': 'nfb_affine_coupling_fwd = rows_warp_kernel<AffineVec<checker>> (split-gather + affine + merge-scatter '
                      '+ per-sample log-det), 3x32x32, B=16384, out-of-place',
            'bound': 'hbm', 'achieved': head['achieved'], 'peak': peak, 'unit': 'GB/s', 'frac': head['achieved'] / peak,
            'traffic': _ncu_traffic('affine_coupling_stream'), 'ms_per_launch': head['ms_per_launch'],
            'alg_bytes_per_launch': alg, 'peak_source': peaks['source'],
            'l2': 'flushed between launches (252 MB write); inputs 604 MB > 126 MB L2',
            'other_splits': {k: {'achieved': v['achieved'], 'frac': v['achieved'] / peak} for k, v in res.items()
                             if k != 'checkerboard 3x32x32'}}


def graph_time_us(fn, per_graph=20, iters=10):
    """median device time of one fn() in us: per_graph launches captured in a CUDA graph, events around replays."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(per_graph):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for e0, e1 in ev:
        e0.record()
        g.replay()
        e1.record()
    torch.cuda.synchronize()
    ts = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    return ts[len(ts) // 2] * 1e3 / per_graph


def dominant_kernel_probe(net, batch, clocks):
    """The kernel with the largest share of the step (profiles/: ~80 % for glow32): the fused ConvNet conditioner.
    It is FP32-FFMA bound (split-precision tensor-core math is future work), so its ceiling is the FP32 pipe:
    148 SMs x 128 FMA/clk x 2 x SM clock.  Timed live, CUDA-graph replay of 20 back-to-back launches."""
    import nfb200
    from nfb200.flows.coupling import AffineCoupling
    seen, out = set(), []
    sm_mhz = (clocks or {}).get('sm_mhz') or 1965.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    for m in net.modules():
        if not isinstance(m, AffineCoupling) or len(m.dims) != 3:
            continue
        C, H, W = m.dims
        h, w = (H // 2, W // 2) if m.mode == nfb200._lib.SPLIT_CHECKER else (H, W)
        key = (m.net.in_channels, m.net.out_channels, h, w)
        if key in seen:
            continue
        seen.add(key)
        z = torch.randn((batch, ) + tuple(m.dims), device='cuda')
        if m.net.forward_from_z(z, m.mode, m.odd) is None:
            continue
        us = graph_time_us(lambda: m.net.forward_from_z(z, m.mode, m.odd))
        mac = h * w * (key[0] * 288 + 4 * 9216 + 32 * key[1])
        tf = 2 * mac * batch / us * 1e-6
        out.append({'kernel': 'nfb_convnet_fwd %dx%d in=%d out=%d' % (h, w, key[0], key[1]), 'us_per_launch': us,
                    'achieved': tf, 'unit': 'TFLOP/s', 'bound': 'fp32-ffma', 'peak': fp32_peak, 'frac': tf / fp32_peak})
    return out


def train_probe(net, dev_ring, batch, world, steps):
    """Training step of main.py:78-92 on the new path: train-mode forward (bijection kernels + cuDNN/cuBLAS conditioner),
    loss = global mean NLL, backward through csrc/backward.cu, flat-bucket gradient all-reduce, Adam -- replayed as CUDA
    graphs (nfb200.parallel.GraphedTrainStep); the eager step is timed next to it.  Runs on deep copies so the
    benchmarked weights are untouched.  Informational: the headline metric is fwd+logdet."""
    import copy
    import nfb200
    from nfb200 import parallel
    out = {'unit': 'samples/s',
           'note': 'train-mode fwd + gradient kernels (bijections and ConvNet conditioner) + flat-bucket all-reduce + Adam'}
    ring = len(dev_ring)
    # eager
    tnet = copy.deepcopy(net).train()
    opt = torch.optim.Adam(tnet.parameters(), lr=1e-4)
    losses = [parallel.train_step(tnet, opt, dev_ring[i % ring]) for i in range(2)]
    torch.cuda.synchronize()
    n0 = nfb200._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(2):
        losses.append(parallel.train_step(tnet, opt, dev_ring[(2 + i) % ring]))
    e1.record()
    torch.cuda.synchronize()
    out['eager'] = {'value': batch * world / (e0.elapsed_time(e1) / 2 * 1e-3), 'ms_per_step': e0.elapsed_time(e1) / 2,
                    'loss_first': losses[0], 'loss_last': losses[-1]}
    out['nfb200_launches_per_step'] = (nfb200._lib.launch_count() - n0) // 2
    del tnet, opt
    # graph replay
    gnet = copy.deepcopy(net).train()
    gopt = torch.optim.Adam(gnet.parameters(), lr=1e-4, capturable=True)
    step = parallel.GraphedTrainStep(gnet, gopt, dev_ring[0])
    first = float(step(dev_ring[0]))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(dev_ring[(1 + i) % ring])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out.update({'value': batch * world / (ms * 1e-3), 'ms_per_step': ms, 'mode': 'cuda-graph replay',
                'loss_first': first, 'loss_last': float(loss)})
    return out


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d.get('bf16_tflops'), 'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'source': 'fallback (B200_PROFILING.md)'}


def run_nfb200(args, rank, world, local_rank):
    import torch.distributed as dist
    import nfb200
    from nfb200 import parallel

    model, dims, datatype, cfg, batch, desc = WORKLOADS[args.workload]
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    D = int(math.prod(dims))

    torch.manual_seed(0)  # identical weights on every rank (replicas)
    net = getattr(nfb200, {'glow': 'Glow', 'flowpp': 'Flowpp', 'realnvp': 'RealNVP'}[model])(
        dims, datatype, types.SimpleNamespace(**cfg)).to(dev).eval()

    # inputs: a ring of distinct batches, larger than L2 in total, pinned on the host for the e2e leg
    bytes_per_batch = batch * D * 4
    ring_n = max(4, min(64, L2_BYTES * 5 // 4 // bytes_per_batch + 1))
    host_ring = [make_inputs(dims, datatype, batch, 1000 * rank + i).pin_memory() for i in range(ring_n)]
    dev_ring = [h.to(dev) for h in host_ring]

    with torch.no_grad():
        net(dev_ring[0])  # ActNorm data-dependent init (modules.py:238-244) + weight caches
        torch.cuda.synchronize()

        # ---- launches per step (eager) -----------------------------------------------------------
        n0 = nfb200._lib.launch_count()
        z, ldj = net(dev_ring[0])
        rows, total = nfb200.gauss_nll(z, ldj)
        launches_per_step = nfb200._lib.launch_count() - n0
        torch.cuda.synchronize()

        # ---- CUDA graphs of one step: `streams` batches in flight, one graph + static buffers per stream ----------
        # N > 1: every lane issues its own all-reduce; 3 lanes is the configuration validated at N = 2 and N = 8
        n_str = max(1, args.streams if world == 1 else min(args.streams, 3))
        lanes = []
        for k in range(n_str):
            st = torch.cuda.Stream()
            x_static = dev_ring[0].clone()
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                for _ in range(2):
                    z, ldj = net(x_static)
                    rows, total = nfb200.gauss_nll(z, ldj)
            st.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=st):
                z, ldj = net(x_static)
                rows, total = nfb200.gauss_nll(z, ldj)
            lanes.append(dict(stream=st, x=x_static, graph=graph, rows=rows, total=total,
                              red=torch.zeros(2, device=dev, dtype=torch.float64),
                              host_total=torch.zeros(2, dtype=torch.float64).pin_memory(),
                              host_rows=torch.zeros(batch, dtype=torch.float32).pin_memory()))
        torch.cuda.synchronize()
        x_static, graph, rows, total = lanes[0]['x'], lanes[0]['graph'], lanes[0]['rows'], lanes[0]['total']

        def step(i, lanes_used, src_ring, read_back):
            ln = lanes[i % lanes_used]
            with torch.cuda.stream(ln['stream']):
                ln['x'].copy_(src_ring[i % ring_n], non_blocking=True)  # device ring (value) or pinned host ring (e2e)
                ln['graph'].replay()
                res = ln['total']
                if world > 1:
                    ln['red'].copy_(ln['total'])
                    dist.all_reduce(ln['red'])
                    res = ln['red']
                if read_back:
                    ln['host_total'].copy_(res, non_blocking=True)
                    ln['host_rows'].copy_(ln['rows'], non_blocking=True)  # per-sample NLL (= -log p(y), main.py:121-124)
            return ln

        def timed(lanes_used, src_ring, read_back):
            """K steps, `lanes_used` batches in flight; device time from CUDA events on the launching stream, which forks
            to / joins the lane streams."""
            cur = torch.cuda.current_stream()
            for i in range(args.warmup):
                step(i, lanes_used, src_ring, read_back)
            for ln in lanes:
                cur.wait_stream(ln['stream'])
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            for ln in lanes[:lanes_used]:
                ln['stream'].wait_event(e0)
            pending = []
            for i in range(args.steps):
                ln = step(args.warmup + i, lanes_used, src_ring, read_back)
                if read_back:  # the caller consumes every result: wait for the step that used this lane before reusing it
                    pending.append(ln)
                    if len(pending) >= lanes_used:
                        p = pending.pop(0)
                        p['stream'].synchronize()
                        _ = p['host_total'][0].item()
            for p in pending:
                p['stream'].synchronize()
                _ = p['host_total'][0].item()
            for ln in lanes[:lanes_used]:
                cur.wait_stream(ln['stream'])
            e1.record(cur)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        ms_single = timed(1, dev_ring, False)  # one batch in flight (also the latency of a step)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        t_wall = time.perf_counter()
        ms = timed(n_str, dev_ring, False)
        wall = time.perf_counter() - t_wall
        if ms * 1e-3 < 1.0 and rank == 0:
            # make sure the clock sampler saw the load: keep the GPU busy a little longer (untimed)
            t_end = time.perf_counter() + 1.0
            while time.perf_counter() < t_end:
                graph.replay()  # local work only: no collective outside the steps every rank executes
            torch.cuda.synchronize()
        clocks = sampler.stop() if rank == 0 else None
        value = batch * world * args.steps / (ms * 1e-3)
        value_single = batch * world * args.steps / (ms_single * 1e-3)

        # ---- end-to-end leg: public API (captured), pinned host inputs, result read back every step ----------------
        e2e_ms = timed(n_str, host_ring, True)
        e2e_value = batch * world * args.steps / (e2e_ms * 1e-3)

        # ---- inverse direction (sampling, main.py:109-116): z -> y, one batch in flight, graph replay ----------------
        inv = None
        try:
            zs = torch.randn_like(dev_ring[0])
            st = lanes[0]['stream']
            with torch.cuda.stream(st):
                for _ in range(2):
                    net.backward(zs)
            st.synchronize()
            ginv = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ginv, stream=st):
                yinv, linv = net.backward(zs)
            with torch.cuda.stream(st):
                for _ in range(3):
                    ginv.replay()
                i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                i0.record(st)
                for _ in range(max(3, args.steps // 2)):
                    ginv.replay()
                i1.record(st)
            st.synchronize()
            inv_ms = i0.elapsed_time(i1) / max(3, args.steps // 2)
            inv = {'value': batch * world / (inv_ms * 1e-3), 'unit': 'samples/s', 'ms_per_step': inv_ms,
                   'note': 'model.backward(z) (inverse + log-det), one batch in flight per GPU'}
        except Exception as e:  # never take the bench line down
            inv = {'error': repr(e)}

        # ---- bits/dim of the (global) batch ring[0] ------------------------------------------------------
        x_static.copy_(dev_ring[0])
        graph.replay()
        torch.cuda.synchronize()
        bpd_local = nfb200.bits_per_dim_from_total(total, D)
        bpd_global = parallel.global_bits_per_dim(total, D)
        torch.cuda.synchronize()

    # ---- training step; every rank takes part (gradient all-reduce) ------------------------------
    train = None
    if not args.no_train and (world == 1 or args.train):
        try:
            train = train_probe(net, dev_ring, batch, world, max(2, min(5, args.steps // 4)))
        except Exception as e:  # never take the bench line down (N > 1: opt-in, a failing rank would stall the others)
            train = {'error': repr(e)}

    if rank != 0:
        return

    peaks = load_peaks()
    roof, dom = None, None
    try:
        roof = coupling_roofline(peaks)
    except Exception as e:  # the roofline probe must never take the bench line down
        roof = {'error': repr(e)}
    try:
        with torch.no_grad():
            dom = dominant_kernel_probe(net, batch, clocks)
    except Exception as e:
        dom = {'error': repr(e)}

    # ---- CPU baseline on the box's host cores: oracle port, bounded sample ---------------------------------
    threads = os.cpu_count() or 1
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    cpu_n = batch if model != 'flowpp' else min(batch, 32)
    if args.workload == 'glow64':
        cpu_n = 32
    xs = host_ring[0][:cpu_n].clone()
    best, bpd_cpu, reps = cpu_forward_timer(args.workload, sd, xs, threads)
    # GPU bits/dim on exactly the CPU's sample
    with torch.no_grad():
        bpd_gpu_sample = net.bits_per_dim(xs.to(dev))
    cpu = {'value': cpu_n / best, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
           'sample': 'best of %d forwards over %d samples of ring[0] (%.2f s each), torch %s CPU, oracle/flow_oracle.py'
                     % (max(reps, 1), cpu_n, best, torch.__version__)}

    out = {
        'metric': 'samples/sec (fwd+logdet)', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': desc, 'batch_per_gpu': batch, 'global_batch': batch * world,
                   'l2': 'inputs rotate over %d distinct batches (%.0f MB > 126 MB L2)' %
                         (ring_n, ring_n * bytes_per_batch / 1e6),
                   'parallelism': 'sample-sharded replicas x%d, all-reduce of (sum NLL, count) only' % world,
                   'batches_in_flight': n_str,
                   'timing': 'CUDA events around K graph replays (%d batches in flight on %d streams), max over ranks' % (n_str, n_str),
                   'wall_s': wall},
        'single_stream': {'value': value_single, 'ms_per_step': ms_single / args.steps,
                          'note': 'one batch in flight: ms_per_step is then the latency of a step'},
        'e2e': {'value': e2e_value, 'unit': 'samples/s', 'ms_per_step': e2e_ms / args.steps,
                'h2d_bytes_per_step': bytes_per_batch, 'd2h_bytes_per_step': 16 + 4 * batch},
        'gpu_launches': launches_per_step * args.steps, 'gpu_launches_per_step': launches_per_step,
        'inverse': inv, 'train_step': train, 'clocks': clocks, 'roofline': roof, 'conditioner_kernels': dom, 'cpu_baseline': cpu,
        'bits_per_dim': {'gpu_global_batch': bpd_global, 'gpu_rank0_batch': bpd_local, 'gpu_on_cpu_sample': bpd_gpu_sample,
                         'cpu_oracle_on_sample': bpd_cpu,
                         'rel_err': abs(bpd_gpu_sample - bpd_cpu) / abs(bpd_cpu), 'tolerance': 1e-5},
    }
    emit(out)


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line goes to the process's real stdout; everything else (NCCL banners, library chatter) was
    re-routed to stderr in main()."""
    line = (json.dumps(obj) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # C libraries (NCCL prints its version banner on stdout) now write to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='nfb200', choices=['nfb200', 'reference'])
    ap.add_argument('--workload', default='glow32', choices=sorted(WORKLOADS))
    ap.add_argument('--streams', type=int, default=5, help='batches in flight (one CUDA graph + stream each)')
    ap.add_argument('--train', action='store_true', help='also time the training step when N > 1 (default: N = 1 only)')
    ap.add_argument('--no-train', action='store_true', help='skip the training-step probe')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'nfb200' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        import datetime
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank),
                                timeout=datetime.timedelta(seconds=90))
    try:
        run_nfb200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
