"""One eager training step (train-mode forward, loss, backward, Adam) of a BASELINE workload, for ncu launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv python profiles/train_step.py glow32 1"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nfb200  # noqa: E402
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else 'glow32'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
model, dims, datatype, cfg, batch, desc = bench.WORKLOADS[wl]
torch.manual_seed(0)
net = getattr(nfb200, {'glow': 'Glow', 'flowpp': 'Flowpp', 'realnvp': 'RealNVP'}[model])(
    dims, datatype, types.SimpleNamespace(**cfg)).cuda().train()
x = bench.make_inputs(dims, datatype, batch, 0).cuda()
opt = torch.optim.Adam(net.parameters(), lr=1e-4)
with torch.no_grad():
    net(x)  # ActNorm init
nfb200.parallel.train_step(net, opt, x)  # warm-up (cuDNN plans, optimizer state)
torch.cuda.synchronize()
n0 = nfb200._lib.launch_count()
torch.cuda.profiler.start()
for _ in range(steps):
    loss = nfb200.parallel.train_step(net, opt, x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('%s: %d steps, loss %.4f, %d libnfb200 launches per step' % (wl, steps, loss, (nfb200._lib.launch_count() - n0) // steps))
