"""Eager (no CUDA graph) steps of a bench workload, for ncu launch lists / full captures.

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv \
        python profiles/glow_step.py [workload] [steps] [throughput]
(`throughput`: nfb200.set_throughput_mode(True), the kernel selection of the bench's 5-lane legs; the profiled range -- cudaProfilerStart/Stop -- is the steps after the initialising forward)
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import nfb200  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else 'glow32'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = bench.WORKLOADS[wl]
model, dims, datatype, cfg, batch = w['model'], w['dims'], w['datatype'], w['cfg'], w['batch']
if len(sys.argv) > 3 and sys.argv[3] == 'throughput':
    nfb200.set_throughput_mode(True)
torch.manual_seed(0)
net = getattr(nfb200, {'glow': 'Glow', 'flowpp': 'Flowpp', 'realnvp': 'RealNVP'}[model])(
    dims, datatype, types.SimpleNamespace(**cfg)).cuda().eval()
x = bench.make_inputs(dims, datatype, batch, 0).cuda()
with torch.no_grad():
    net(x)  # ActNorm init + weight packing (skipped with ncu -s)
    torch.cuda.synchronize()
    n0 = nfb200._lib.launch_count()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(steps):
        z, ldj = net(x)
        rows, total = nfb200.gauss_nll(z, ldj)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print('launches per step', (nfb200._lib.launch_count() - n0) // steps, 'bpd', nfb200.bits_per_dim_from_total(total, x[0].numel()))
