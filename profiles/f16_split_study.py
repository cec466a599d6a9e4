"""CPU study behind the operand format of the tensor-core conditioner: a 6-layer 32-channel 3x3 stack evaluated
with exact (fp64) accumulation of the products each scheme feeds the tensor core -- fp32, 3xTF32 (hi = tf32_rn(x)), the FP16
split (hi = fp16(x), lo = fp16(2^10 (x - hi))) and single-pass TF32 -- against the fp64 result, at three input scales.
    python profiles/f16_split_study.py  ->  profiles/r02b_f16_split_study.txt"""
import torch, math
torch.manual_seed(0)
def tf32_rn(x):
    xi = x.view(torch.int32)
    r = ((xi + 0x0fff + ((xi >> 13) & 1)) & ~0x1fff)
    return r.view(torch.float32)
def split_tf32(x):
    hi = tf32_rn(x); lo = tf32_rn(x - hi); return hi, lo
S = 10
def split_f16(x):
    hi = x.clamp(-65504, 65504).half()
    lo = ((x - hi.float()) * 2.0**S).clamp(-65504, 65504).half()
    return hi.float(), lo.float()
import torch.nn.functional as F
def conv64(a, w): return F.conv2d(a.double(), w.double(), padding=w.shape[-1]//2)
def run(kind, x, Ws):
    a = x
    for i, w in enumerate(Ws):
        if kind == 'f64':
            o = conv64(a, w)
        elif kind == 'f32':
            o = F.conv2d(a.float(), w.float(), padding=w.shape[-1]//2).double()
        elif kind == 'tf32x3':
            ah, al = split_tf32(a.float()); wh, wl = split_tf32(w.float())
            o = (conv64(ah, wh).float() + (conv64(ah, wl) + conv64(al, wh)).float()).double()
        elif kind == 'f16x3':
            ah, al = split_f16(a.float()); wh, wl = split_f16(w.float())
            o = (conv64(ah, wh).float() + ((conv64(ah, wl) + conv64(al, wh)).float() * 2.0**-S)).double()
        elif kind == 'tf32x1':
            ah, _ = split_tf32(a.float()); wh, _ = split_tf32(w.float())
            o = conv64(ah, wh)
        if i < len(Ws) - 1:
            o = torch.relu(o)
        a = o if kind == 'f64' else o.float().double()
    return a
for scale in (1.0, 1e-3, 300.0):
    x = torch.randn(4, 32, 16, 16) * scale
    Ws = [torch.randn(32, 32, 3, 3) / math.sqrt(288) for _ in range(5)] + [torch.randn(12, 32, 1, 1) / math.sqrt(32)]
    ref = run('f64', x.double(), Ws)
    for kind in ('f32', 'tf32x3', 'f16x3', 'tf32x1'):
        o = run(kind, x.double(), Ws)
        print('scale %g %-7s max err / max|ref| = %.3e' % (scale, kind, float((o - ref).abs().max() / ref.abs().max())))
