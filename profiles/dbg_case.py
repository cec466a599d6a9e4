import sys, torch
sys.path.insert(0, '/root/repo')
import nfb200
torch.set_grad_enabled(False)
torch.manual_seed(7)
dims, masking = (192, 8, 8), 'checkerboard'
cpl = nfb200.flows.AffineCoupling(dims, masking=masking, odd=False).cuda().eval()
for B in (1, 3, 19, 256):
    x = torch.randn((B,) + dims, device='cuda'); l0 = torch.randn(B, device='cuda')
    z1, l1 = cpl(x, l0.clone())
    torch.cuda.synchronize()
    print('B', B, 'ok', float(z1.abs().max()))
