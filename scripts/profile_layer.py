"""Run one Hebbian conv layer a few times (for ncu).  usage: profile_layer.py Cin Cout k H W B prec [nd D]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200'))
import torch, hebb
Cin, Cout, k, H, W, B = map(int, sys.argv[1:7]); prec = sys.argv[7]
D = int(sys.argv[8]) if len(sys.argv) > 8 else 0
cls = hebb.HebbianConv3d if D else hebb.HebbianConv2d
layer = cls(Cin, Cout, k, padding=k // 2, bias=False, k=50., alpha=1.)
layer.prec = prec
layer = layer.cuda().train()
x = torch.randn(B, Cin, *( (D, H, W) if D else (H, W) ), device='cuda')
for _ in range(3):
    layer(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); layer(x); e1.record(); torch.cuda.synchronize()
print('layer ms', e0.elapsed_time(e1))
