#!/usr/bin/env python
"""bench.py -- samples/sec of the flow forward + log-det + NLL hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl nfb200|reference] [--workload NAME]

One "step" = one eval-mode pass of the whole stack over one batch of synthetic inputs:
``model.forward(x) -> (z, log_df_dz)`` followed by the Gaussian NLL reduction; for N > 1 the (sum NLL, count) payloads
of all steps of the timed window are all-reduced ONCE, inside the timed region -- the only collective.  Headline
workload: BASELINE.json configs[1], Glow K=32 L=3 on 32x32x3, batch 256 per GPU (weak scaling).

Prints ONE JSON line (rank 0).  `value` = whole-job samples/s with inputs resident in HBM (CUDA-graph replay, CUDA-event
timing, max over ranks); `e2e` = the same through the public nn.Module API with pinned HOST buffers (H2D copy of x and
D2H read of the NLL inside the timed region); `roofline` = the dominant hand-written kernel (tensor-core conditioner +
coupling), timed live; `roofline_hbm` = the streaming affine-coupling kernel the north-star target names;
`configs` = every other BASELINE.json config (value, e2e, bits/dim parity), cfg 3 / 4 with their global batch split over
the N ranks (strong scaling); `cpu_baseline` = the reference's CPU path on the box's host cores (the reference itself
from oracle/_ref when it travelled, else the oracle port).  `--impl reference` times that CPU path alone.
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

# batch: samples per GPU of the weak-scaling line; strong: the global batch BASELINE.json fixes for the config (split
# over the ranks), None = no multi-GPU split named
WORKLOADS = {
    'glow32': dict(model='glow', dims=(3, 32, 32), datatype='image', cfg=dict(layers=32, mixtures=4), batch=256, strong=None,
                   desc='Glow K=32 L=3 on 32x32x3 synthetic images, batch 256 per GPU (BASELINE.json configs[1])'),
    'flowpp32': dict(model='flowpp', dims=(3, 32, 32), datatype='image', cfg=dict(layers=32, mixtures=8), batch=256,
                     strong=None, desc='Flow++ logistic-mixture coupling on 32x32x3 synthetic, batch 256 per GPU (configs[2])'),
    'realnvp64_rqs': dict(model='realnvp', dims=(64, ), datatype=None, cfg=dict(layers=8, mixtures=4, coupling='rqs'),
                          batch=65536, strong=65536,
                          desc='RealNVP 8 RQ-spline couplings on 64-dim synthetic tabular, global batch 65536 (configs[3])'),
    'realnvp64': dict(model='realnvp', dims=(64, ), datatype=None, cfg=dict(layers=8, mixtures=4), batch=65536, strong=65536,
                      desc='RealNVP 8 affine couplings on 64-dim synthetic tabular, batch 65536 (configs[3], affine proxy)'),
    'realnvp2': dict(model='realnvp', dims=(2, ), datatype=None, cfg=dict(layers=6, mixtures=4), batch=512, strong=None,
                     desc='RealNVP 6 affine couplings on 2D, batch 512 (configs[0])'),
    'glow64': dict(model='glow', dims=(3, 64, 64), datatype='image', cfg=dict(layers=48, mixtures=4), batch=256, strong=2048,
                   desc='Glow K=48 L=4 on 64x64x3 synthetic, global batch 2048 sharded over the GPUs (configs[4])'),
}
MODEL_CLASS = {'glow': 'Glow', 'flowpp': 'Flowpp', 'realnvp': 'RealNVP'}


def make_inputs(dims, datatype, batch, seed):
    g = torch.Generator().manual_seed(seed)
    if datatype == 'image':
        return torch.rand((batch, ) + tuple(dims), generator=g)
    return torch.randn((batch, ) + tuple(dims), generator=g)


def config_of(wl, world, scaling):
    """The `config` object: identical in the nfb200 arm and the reference arm for the same flags."""
    W = WORKLOADS[wl]
    if scaling == 'strong':
        per, glob = W['strong'] // world, W['strong']
    else:
        per, glob = W['batch'], W['batch'] * world
    return {'workload': W['desc'], 'batch_per_gpu': per, 'global_batch': glob}


# ---------------------------------------------------------------------------------------------------------------------
# CPU side: the reference itself (oracle/_ref, copied there by __graft_entry__.build() where /root/reference exists; it is
# git-ignored but travels with the snapshot) or the oracle port.  Test / baseline infrastructure only.
# ---------------------------------------------------------------------------------------------------------------------
def cpu_path(wl):
    """-> (forward(sd_or_None, x) -> (z, ldj), kind, description).  sd: state dict to load (parity), None: own init."""
    W = WORKLOADS[wl]
    cfg = types.SimpleNamespace(**W['cfg'])
    ref_dir = os.path.join(ROOT, 'oracle', '_ref')
    if W['cfg'].get('coupling') is None and os.path.isdir(os.path.join(ref_dir, 'flows')):
        try:
            if ref_dir not in sys.path:
                sys.path.insert(0, ref_dir)
            import warnings
            warnings.filterwarnings('ignore')
            import flows as ref_flows  # the unmodified reference package
            torch.manual_seed(0)
            net = getattr(ref_flows, MODEL_CLASS[W['model']])(W['dims'], W['datatype'], cfg).eval()

            def forward(sd, x):
                if sd is not None:
                    net.load_state_dict(sd, strict=True)
                    for m in net.modules():
                        if hasattr(m, 'initialized'):
                            m.initialized = True  # not part of the state dict (modules.py:235)
                return net(x)
            return forward, 'reference', 'tatsy/normalizing-flows-pytorch flows.%s (oracle/_ref), torch %s CPU' % (
                MODEL_CLASS[W['model']], torch.__version__)
        except Exception as e:  # fall through to the port
            sys.stderr.write('reference import failed (%r); using the oracle port\n' % (e, ))
    from oracle import flow_oracle as O
    spec = O.stack_spec(W['model'], W['dims'], W['datatype'], W['cfg']['layers'], W['cfg'].get('mixtures', 4),
                        W['cfg'].get('coupling'))
    state = {}

    def forward(sd, x):
        if sd is not None:
            state['sd'] = sd
        if 'sd' not in state:  # random init of the right architecture (only reached when the reference did not travel)
            state['sd'] = port_state_dict(wl)
        return O.stack_forward(spec, state['sd'], x)
    return forward, 'port', 'oracle/flow_oracle.py (port of the reference CPU path), torch %s CPU' % torch.__version__


def cpu_bits_per_dim(z, ldj):
    from oracle import flow_oracle as O
    return O.bits_per_dim(z, ldj)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on all host threads, same workload, metric
    and config object; each step a bounded sample of the batch so the run ends within minutes.  Rank 0 only."""
    if rank != 0:
        return
    W = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    forward, kind, what = cpu_path(args.workload)
    sd = None
    probe_n = min(W['batch'], 16)
    xp = make_inputs(W['dims'], W['datatype'], probe_n, 123)
    with torch.no_grad():
        forward(sd, xp)  # warm-up; the reference runs its ActNorm data-dependent init here
        t0 = time.perf_counter()
        forward(None, xp)
        per_sample = (time.perf_counter() - t0) / probe_n
    n_steps = args.steps + args.warmup
    sample = int(max(1, min(W['batch'], 120.0 / max(per_sample * n_steps, 1e-9))))
    x = make_inputs(W['dims'], W['datatype'], sample, 0)
    with torch.no_grad():
        for _ in range(args.warmup):
            forward(None, x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            forward(None, x)
        dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    emit({
        'impl': 'reference', 'metric': 'samples/sec (fwd+logdet)', 'value': value, 'unit': 'samples/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_of(args.workload, world, args.scaling),
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': threads, 'kind': kind,
                         'sample': '%d samples per step of the %d-sample batch; %s' % (sample, W['batch'], what)},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    })


def port_state_dict(wl):
    """Fallback only (oracle/_ref absent, or the RQ-spline workload the reference does not have): random-init CPU state
    dict of the workload's architecture for the oracle port, taken from the package's reference-compatible module tree
    (parameter shapes only: no kernels, no CUDA)."""
    import nfb200
    W = WORKLOADS[wl]
    torch.manual_seed(0)
    net = getattr(nfb200, MODEL_CLASS[W['model']])(W['dims'], W['datatype'], types.SimpleNamespace(**W['cfg']))
    return {k: v.clone() for k, v in net.state_dict().items()}


# ---------------------------------------------------------------------------------------------------------------------
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
                                       '--format=csv,noheader,nounits', '-lms', '50'], stdout=self.f,
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


def time_kernel_stream(fn, iters, flush=None):
    """Device time of fn() in ms (mean, median, best), CUDA events on the current stream, optional L2 flush between calls."""
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


def _ncu(key, field):
    """A number from the committed ncu --set full summaries (profiles/ncu_summary.json)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_summary.json')) as f:
            return json.load(f).get(key, {}).get(field)
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
            'traffic': _ncu('affine_coupling_stream', 'dram_bytes_per_launch'), 'ms_per_launch': head['ms_per_launch'],
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


def dominant_kernel_probe(net, batch, peaks):
    """The kernels with the largest share of the step (profiles/r02_launches_glow32.md: ~85 %): the tensor-core ConvNet
    conditioner with the affine coupling as its epilogue (nfb_convnet_affine_fwd), one per conditioner shape.  Timed live,
    CUDA-graph replay of 20 back-to-back launches on inputs of the benchmark's size.  `achieved` counts the ALGORITHMIC
    flops of the fp32 convolution stack (2 x MAC) -- the kernel issues 3 half-precision products per MAC
    (error-compensated FP16 split: hi*hi + hi*lo + lo*hi), so its ceiling on this formulation is peak_f16 / 3 (fp16 and bf16
    share one tensor-pipe rate; MEASURED_PEAKS.json has the dense bf16 number)."""
    import nfb200
    from nfb200.flows.coupling import AffineCoupling
    seen, out, count = set(), [], {}
    peak = peaks.get('bf16_tflops') or 1590.0
    for m in net.modules():
        if isinstance(m, AffineCoupling) and len(m.dims) == 3:
            C, H, W = m.dims
            h, w = (H // 2, W // 2) if m.mode == nfb200._lib.SPLIT_CHECKER else (H, W)
            key = (m.net.in_channels, m.net.out_channels, h, w)
            count[key] = count.get(key, 0) + 1
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
        ldj = torch.zeros(batch, device='cuda')
        if m.forward_fused(z, ldj, inplace=True) is None:
            continue
        us = graph_time_us(lambda: m.forward_fused(z, ldj, inplace=True))
        mac = h * w * (key[0] * 288 + 4 * 9216 + 32 * key[1])
        tf = 2 * mac * batch / us * 1e-6
        out.append({'kernel': 'nfb_convnet_affine_fwd %dx%d in=%d out=%d (convnet_tc_kernel: tcgen05 FP16-split x3 + coupling epilogue)'
                              % (h, w, key[0], key[1]),
                    'bound': 'tensor', 'us_per_launch': us, 'launches_per_step': count[key], 'achieved': tf, 'peak': peak, 'unit': 'TFLOP/s', 'frac': tf / peak,
                    'alg_flops_per_launch': 2 * mac * batch, 'frac_of_split_ceiling': tf / (peak / 3.0),
                    'traffic': _ncu('convnet_tc_%dx%d' % (h, w), 'dram_bytes_per_launch'),
                    'tensor_pipe_active_pct': _ncu('convnet_tc_%dx%d' % (h, w), 'tensor_pipe_active_pct'),
                    'peak_source': peaks['source'] + ' (dense bf16)'})
    out.sort(key=lambda r: -r['us_per_launch'] * r['launches_per_step'])  # largest share of the step first
    return out


def train_probe(net, dev_ring, batch, world, steps):
    """Training step of main.py:78-92 on the new path: train-mode forward, loss = global mean NLL, backward through
    csrc/backward.cu + conditioner_train.cu, flat-bucket gradient all-reduce, Adam -- replayed as CUDA graphs
    (nfb200.parallel.GraphedTrainStep); the eager step is timed next to it.  Runs on deep copies so the benchmarked
    weights are untouched.  Informational: the headline metric is fwd+logdet."""
    import copy
    import nfb200
    from nfb200 import parallel
    out = {'unit': 'samples/s',
           'note': 'train-mode fwd + gradient kernels (bijections and ConvNet conditioner) + flat-bucket all-reduce + Adam'}
    ring = len(dev_ring)
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


# ---------------------------------------------------------------------------------------------------------------------
class Harness:
    """One workload on this rank: model, a ring of distinct input batches (larger than L2 in total; pinned on the host for
    the e2e leg) and `n_lanes` captured CUDA graphs of one step, each with its own stream and static buffers."""

    def __init__(self, wl, batch, rank, world, dev, n_lanes, ring_cap=64):
        import nfb200
        W = WORKLOADS[wl]
        self.W, self.wl, self.batch, self.rank, self.world, self.dev = W, wl, batch, rank, world, dev
        self.D = int(math.prod(W['dims']))
        torch.manual_seed(0)  # identical weights on every rank (replicas)
        self.net = getattr(nfb200, MODEL_CLASS[W['model']])(W['dims'], W['datatype'],
                                                            types.SimpleNamespace(**W['cfg'])).to(dev).eval()
        self.bytes_per_batch = batch * self.D * 4
        self.ring_n = max(4, min(ring_cap, L2_BYTES * 5 // 4 // self.bytes_per_batch + 1))
        self.host_ring = [make_inputs(W['dims'], W['datatype'], batch, 1000 * rank + i).pin_memory() for i in range(self.ring_n)]
        self.dev_ring = [h.to(dev) for h in self.host_ring]
        with torch.no_grad():
            self.net(self.dev_ring[0])  # ActNorm data-dependent init (modules.py:238-244) + weight caches
            torch.cuda.synchronize()
            n0 = nfb200._lib.launch_count()
            z, ldj = self.net(self.dev_ring[0])
            nfb200.gauss_nll(z, ldj)
            self.launches_per_step = nfb200._lib.launch_count() - n0
            torch.cuda.synchronize()
            # several batches in flight: two tiles per CTA on the 8x8 / 4x4 conditioner maps (nfb200.set_throughput_mode);
            # the extra last lane is captured in latency mode and serves the single-stream leg
            self.lanes = []
            for k in range(n_lanes + (1 if n_lanes > 1 else 0)):
                nfb200.set_throughput_mode(n_lanes > 1 and k < n_lanes)
                st = torch.cuda.Stream()
                x_static = self.dev_ring[0].clone()
                st.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(st):
                    for _ in range(2):
                        z, ldj = self.net(x_static)
                        rows, total = nfb200.gauss_nll(z, ldj)
                st.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=st):
                    z, ldj = self.net(x_static)
                    rows, total = nfb200.gauss_nll(z, ldj)
                self.lanes.append(dict(stream=st, x=x_static, graph=graph, rows=rows, total=total,
                                       host_total=torch.zeros(2, dtype=torch.float64).pin_memory(),
                                       host_rows=torch.zeros(batch, dtype=torch.float32).pin_memory()))
            nfb200.set_throughput_mode(False)
            self.latency_lane = self.lanes.pop() if n_lanes > 1 else self.lanes[0]
            torch.cuda.synchronize()

    def timed(self, steps, warmup, lanes_used, host_inputs, read_back):
        """`steps` steps with `lanes_used` batches in flight.  Device time from CUDA events on the launching stream, which
        forks to / joins the lane streams.  N > 1: every step's (sum NLL, count) lands in one device array that is
        all-reduced ONCE after the last step, inside the timed region; result = max over ranks."""
        import torch.distributed as dist
        src_ring = self.host_ring if host_inputs else self.dev_ring
        lanes, world = (self.lanes if lanes_used > 1 else [self.latency_lane]), self.world
        totals = torch.zeros(steps + warmup, 2, device=self.dev, dtype=torch.float64)
        host_totals = torch.zeros(steps + warmup, 2, dtype=torch.float64).pin_memory()

        def step(i):
            ln = lanes[i % lanes_used]
            with torch.cuda.stream(ln['stream']):
                ln['x'].copy_(src_ring[i % self.ring_n], non_blocking=True)  # device ring (value) or pinned host ring (e2e)
                ln['graph'].replay()
                totals[i].copy_(ln['total'], non_blocking=True)
                if read_back:
                    ln['host_total'].copy_(ln['total'], non_blocking=True)
                    ln['host_rows'].copy_(ln['rows'], non_blocking=True)  # per-sample NLL (= -log p(y), main.py:121-124)
            return ln

        with torch.no_grad():
            cur = torch.cuda.current_stream()
            for i in range(warmup):
                step(i)
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
            for i in range(steps):
                ln = step(warmup + i)
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
            if world > 1:
                dist.all_reduce(totals)  # the only collective: K x (sum NLL, count) in one launch-latency-bound all-reduce
            if read_back:
                host_totals.copy_(totals, non_blocking=True)
            e1.record(cur)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=self.dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

    def bits_per_dim_ring0(self):
        import nfb200
        from nfb200 import parallel
        ln = self.latency_lane
        with torch.no_grad():
            ln['x'].copy_(self.dev_ring[0])
            ln['graph'].replay()
            torch.cuda.synchronize()
            local = nfb200.bits_per_dim_from_total(ln['total'], self.D)
            glob = parallel.global_bits_per_dim(ln['total'], self.D)
            torch.cuda.synchronize()
        return local, glob

    def cpu_check(self, cpu_n, budget_s=20.0):
        """The reference's CPU path on the first cpu_n samples of ring[0] with THIS model's weights: throughput on the
        host cores (best of <= 3) and bits/dim parity of the GPU path on exactly that sample."""
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        forward, kind, what = cpu_path(self.wl)
        sd = {k: v.detach().cpu().clone() for k, v in self.net.state_dict().items()}
        xs = self.host_ring[0][:cpu_n].clone()
        with torch.no_grad():
            t0 = time.perf_counter()
            z, ldj = forward(sd, xs)
            best, reps = time.perf_counter() - t0, 0
            stop = time.perf_counter() + budget_s
            while reps < 3 and time.perf_counter() + best < stop:
                t0 = time.perf_counter()
                z, ldj = forward(None, xs)
                best = min(best, time.perf_counter() - t0)
                reps += 1
            bpd_cpu = cpu_bits_per_dim(z, ldj)
            bpd_gpu = self.net.bits_per_dim(xs.to(self.dev))
        return ({'value': cpu_n / best, 'unit': 'samples/s', 'cores': threads, 'kind': kind,
                 'sample': 'best of %d forwards over %d samples of ring[0] (%.2f s each); %s' % (max(reps, 1), cpu_n, best, what)},
                {'gpu_on_cpu_sample': bpd_gpu, 'cpu_on_sample': bpd_cpu, 'rel_err': abs(bpd_gpu - bpd_cpu) / abs(bpd_cpu),
                 'tolerance': 1e-5})


def library_calls():
    """conditioner calls that left libnfb200 for cuDNN / cuBLAS torch ops since import (0 on the benchmark paths)."""
    try:
        import nfb200
        return int(nfb200.library_path_calls())
    except Exception:
        return None


def run_config(args, wl, scaling, rank, world, dev, n_lanes, light):
    """One BASELINE config -> dict.  light: fewer steps, no single-stream leg (the secondary configs)."""
    W = WORKLOADS[wl]
    cfg = config_of(wl, world, scaling)
    batch = cfg['batch_per_gpu']
    steps = max(3, args.steps // 4) if light else args.steps
    warmup = 2 if light else args.warmup
    lib0 = library_calls()
    H = Harness(wl, batch, rank, world, dev, n_lanes, ring_cap=16 if light else 64)
    ms = H.timed(steps, warmup, n_lanes, False, False)
    e2e_ms = H.timed(steps, warmup, n_lanes, True, True)
    out = {'config': cfg, 'scaling': scaling, 'value': batch * world * steps / (ms * 1e-3), 'unit': 'samples/s',
           'ms_per_step': ms / steps, 'steps': steps, 'batches_in_flight': n_lanes,
           'e2e': {'value': batch * world * steps / (e2e_ms * 1e-3), 'unit': 'samples/s', 'ms_per_step': e2e_ms / steps,
                   'h2d_bytes_per_step': H.bytes_per_batch, 'd2h_bytes_per_step': 16 + 4 * batch},
           'gpu_launches_per_step': H.launches_per_step}
    if lib0 is not None:
        out['library_launches'] = library_calls() - lib0  # conditioner calls that fell back to cuDNN / cuBLAS torch ops
    local, glob = H.bits_per_dim_ring0()
    out['bits_per_dim_global'] = glob
    if rank == 0 and world == 1:
        cpu_n = {'flowpp32': 8, 'glow64': 8, 'glow32': 32}.get(wl, min(batch, 65536))
        try:
            cpu, par = H.cpu_check(min(cpu_n, batch), budget_s=6.0)
            out['cpu_baseline'], out['bits_per_dim'] = cpu, par
        except Exception as e:
            out['bits_per_dim'] = {'error': repr(e)}
    return out, H


def run_nfb200(args, rank, world, local_rank):
    import nfb200
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    wl, scaling = args.workload, args.scaling
    W = WORKLOADS[wl]
    cfg = config_of(wl, world, scaling)
    batch = cfg['batch_per_gpu']
    n_lanes = max(1, args.streams)  # the same number of batches in flight at every N

    lib0 = library_calls()
    H = Harness(wl, batch, rank, world, dev, n_lanes)
    ms_single = H.timed(args.steps, args.warmup, 1, False, False)  # one batch in flight (also the latency of a step)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_wall = time.perf_counter()
    ms = H.timed(args.steps, args.warmup, n_lanes, False, False)
    wall = time.perf_counter() - t_wall
    if ms * 1e-3 < 1.0 and rank == 0:
        # make sure the clock sampler saw load: keep the GPU busy a little longer (untimed, local work only)
        t_end = time.perf_counter() + 1.0
        with torch.no_grad():
            while time.perf_counter() < t_end:
                H.lanes[0]['graph'].replay()  # local work only
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    value = batch * world * args.steps / (ms * 1e-3)
    value_single = batch * world * args.steps / (ms_single * 1e-3)
    e2e_ms = H.timed(args.steps, args.warmup, n_lanes, True, True)
    e2e_value = batch * world * args.steps / (e2e_ms * 1e-3)
    lib_calls = None if lib0 is None else library_calls() - lib0

    # ---- inverse direction (sampling, main.py:109-116): z -> y, one batch in flight, graph replay ----------------
    try:
        with torch.no_grad():
            zs = torch.randn_like(H.dev_ring[0])
            st = H.latency_lane['stream']
            with torch.cuda.stream(st):
                for _ in range(2):
                    H.net.backward(zs)
            st.synchronize()
            ginv = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ginv, stream=st):
                H.net.backward(zs)
            n_inv = max(3, args.steps // 2)
            with torch.cuda.stream(st):
                for _ in range(3):
                    ginv.replay()
                i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                i0.record(st)
                for _ in range(n_inv):
                    ginv.replay()
                i1.record(st)
            st.synchronize()
            inv_ms = i0.elapsed_time(i1) / n_inv
        inv = {'value': batch * world / (inv_ms * 1e-3), 'unit': 'samples/s', 'ms_per_step': inv_ms,
               'note': 'model.backward(z) (inverse + log-det), one batch in flight per GPU'}
    except Exception as e:  # never take the bench line down
        inv = {'error': repr(e)}

    bpd_local, bpd_global = H.bits_per_dim_ring0()
    launches_per_step, ring_n, bytes_per_batch = H.launches_per_step, H.ring_n, H.bytes_per_batch

    # ---- training step; every rank takes part (gradient all-reduce) ------------------------------
    train = None
    if not args.no_train and (world == 1 or args.train):
        try:
            train = train_probe(H.net, H.dev_ring, batch, world, max(2, min(5, args.steps // 4)))
        except Exception as e:  # never take the bench line down (N > 1: opt-in, a failing rank would stall the others)
            train = {'error': repr(e)}

    # ---- rank-0-only probes of the headline workload -------------------------------------------------------------------
    roof_hbm = dom = cpu = parity = None
    if rank == 0:
        peaks = load_peaks()
        try:
            roof_hbm = coupling_roofline(peaks)
        except Exception as e:  # the roofline probe must never take the bench line down
            roof_hbm = {'error': repr(e)}
        try:
            with torch.no_grad():
                dom = dominant_kernel_probe(H.net, batch, peaks)
        except Exception as e:
            dom = [{'error': repr(e)}]
        cpu_n = {'flowpp32': 32, 'glow64': 32}.get(wl, batch)
        cpu, parity = H.cpu_check(min(cpu_n, batch))
        parity.update({'gpu_global_batch': bpd_global, 'gpu_rank0_batch': bpd_local})
    net_main = H.net
    del H
    torch.cuda.empty_cache()

    # ---- every other BASELINE config (all ranks take part: each window ends with the all-reduce) --------------------
    configs = {}
    if not args.no_configs and wl == 'glow32':
        plan = [('flowpp32', 'weak'), ('realnvp64_rqs', 'strong'), ('glow64', 'strong')]
        if world == 1:
            plan.insert(0, ('realnvp2', 'weak'))
        for name, sc in plan:
            try:
                res, Hc = run_config(args, name, sc, rank, world, dev, n_lanes, True)
                configs[name] = res
                del Hc
            except Exception as e:
                configs[name] = {'error': repr(e)}
            torch.cuda.empty_cache()
    del net_main

    if rank != 0:
        return
    roofline = dom[0] if dom and 'error' not in dom[0] else roof_hbm
    out = {
        'metric': 'samples/sec (fwd+logdet)', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
        'run': {'l2': 'inputs rotate over %d distinct batches (%.0f MB > 126 MB L2)' % (ring_n, ring_n * bytes_per_batch / 1e6),
                'parallelism': 'sample-sharded replicas x%d; one all-reduce of K x (sum NLL, count) per timed window' % world,
                'batches_in_flight': n_lanes,
                'kernel_selection': 'throughput mode (nfb200.set_throughput_mode: two units in flight per CTA on every conditioner map size) '
                                    'for the %d-lane legs; latency mode for single_stream and inverse' % n_lanes,
                'timing': 'CUDA events around K graph replays (%d batches in flight on %d streams), max over ranks' % (n_lanes, n_lanes),
                'wall_s': wall},
        'single_stream': {'value': value_single, 'ms_per_step': ms_single / args.steps,
                          'note': 'one batch in flight: ms_per_step is then the latency of a step'},
        'e2e': {'value': e2e_value, 'unit': 'samples/s', 'ms_per_step': e2e_ms / args.steps,
                'h2d_bytes_per_step': bytes_per_batch, 'd2h_bytes_per_step': 16 + 4 * batch},
        'gpu_launches': launches_per_step * args.steps, 'gpu_launches_per_step': launches_per_step,
        'library_launches': lib_calls,
        'inverse': inv, 'train_step': train, 'clocks': clocks, 'roofline': roofline, 'roofline_hbm': roof_hbm,
        'conditioner_kernels': dom, 'cpu_baseline': cpu, 'bits_per_dim': parity, 'configs': configs,
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
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='strong: the global batch BASELINE.json names for the workload is split over the ranks')
    ap.add_argument('--streams', type=int, default=5, help='batches in flight (one CUDA graph + stream each)')
    ap.add_argument('--train', action='store_true', help='also time the training step when N > 1 (default: N = 1 only)')
    ap.add_argument('--no-train', action='store_true', help='skip the training-step probe')
    ap.add_argument('--no-configs', action='store_true', help='skip the block with the other BASELINE configs')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'nfb200' else args.warmup
    if args.scaling == 'strong' and WORKLOADS[args.workload]['strong'] is None:
        ap.error('BASELINE.json names no global batch for %s' % args.workload)

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
                                timeout=datetime.timedelta(seconds=180))
    try:
        run_nfb200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
