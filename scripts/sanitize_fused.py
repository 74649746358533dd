"""Small self-checking invocations of every fused-kernel variant (forward + update, forward only, patch gather, 1x1,
weight-gradient mode in both layouts, winner fix-up) against the fp32 CUDA-core path / the fp64 gradient.  Written as the
target of `compute-sanitizer --tool memcheck|synccheck|racecheck python scripts/sanitize_fused.py`; the sanitizer is
closed on the GPU pool this was developed on, so there it only runs plain (every wait in these kernels is bounded and
traps with a code readable through hebb_watchdog_code)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200'))
import torch
import hebb
from hebb import _native as N

dev = torch.device('cuda', 0)
torch.manual_seed(0)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


worst = 0.0
for Cin, Cout, k, B, H, W in ((16, 16, 3, 2, 24, 28), (32, 32, 3, 2, 20, 24), (16, 32, 3, 1, 40, 36), (32, 16, 1, 2, 16, 20), (3, 16, 3, 2, 32, 36)):
    x = torch.randn(B, Cin, H, W, device=dev)
    out = {}
    for prec in ('fp32', 'bf16x3'):
        torch.manual_seed(1)
        layer = hebb.HebbianConv2d(Cin, Cout, k, padding=k // 2, bias=True, k=20., alpha=1.)
        layer.prec = prec
        layer.record_winners = True
        layer = layer.to(dev).train()
        if prec == 'bf16x3':
            assert N.layer_path(layer._desc(x.shape, True), N.PREC_BF16X3, N.F_UPDATE | N.F_WNRM) == N.PATH_FUSED, (Cin, Cout, k)
        y = layer(x)
        layer.eval()
        with torch.no_grad():
            y2 = layer(x)                      # forward only
        out[prec] = (y, layer.delta_w.clone(), layer.winners.clone(), y2)
    e = max(rel(out['bf16x3'][0], out['fp32'][0]), rel(out['bf16x3'][1], out['fp32'][1]), rel(out['bf16x3'][3], out['fp32'][3]))
    nbad = int((out['bf16x3'][2] != out['fp32'][2]).sum())
    print(f'fused {Cin}->{Cout} k{k}: err {e:.2e} winner mismatches {nbad}', flush=True)
    worst = max(worst, e)
    assert e < 1e-4 and nbad <= 1

for Cin, Cout, gyc, k, B, H, W in ((16, 64, 64, 3, 2, 24, 28), (64, 32, 32, 3, 1, 20, 24), (32, 16, 2, 3, 2, 16, 20), (16, 16, 16, 1, 2, 12, 16)):
    x = torch.randn(B, Cin, H, W, device=dev)
    gy = torch.randn(B, gyc, H, W, device=dev)
    desc = N.make_desc(2, B, Cin, Cout, (H, W), (k, k), (1, 1), (k // 2, k // 2), (k // 2, k // 2), False)
    assert N.wgrad_path(desc, N.PREC_BF16X3) == N.PATH_FUSED
    ref = torch.nn.grad.conv2d_weight(x.double(), (gyc, Cin, k, k), gy.double(), padding=k // 2)
    for cl in (False, True):
        xs = x.contiguous(memory_format=torch.channels_last) if cl else x
        gs = gy.contiguous(memory_format=torch.channels_last) if cl else gy
        gw = N.conv_wgrad(desc, xs, gs, N.PREC_BF16X3, gy_channels=gyc, channels_last=cl)
        e = rel(gw[:gyc].reshape(gyc, Cin, k, k), ref)
        print(f'wgrad {Cin}->{gyc} k{k} {"nhwc" if cl else "nchw"}: err {e:.2e}', flush=True)
        worst = max(worst, e)
        assert e < 1e-4
torch.cuda.synchronize()
print(f'sanitize_fused ok, worst error {worst:.2e}')
