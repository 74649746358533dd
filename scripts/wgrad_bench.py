"""Time hebb_conv_wgrad (fused kernel, weight-gradient mode) against cuDNN's weight gradient on the 2-D head's layers.
usage: wgrad_bench.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200'))
import torch
from hebb import _native as N

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for Cin, Cout, gyc in ((16, 64, 64), (64, 32, 32), (32, 16, 2), (32, 32, 32), (16, 16, 16)):
    x = torch.randn(B, Cin, 256, 256, device=dev)
    gy = torch.randn(B, gyc, 256, 256, device=dev)
    desc = N.make_desc(2, B, Cin, Cout, (256, 256), (3, 3), (1, 1), (1, 1), (1, 1), False)
    row = [f'{Cin}->{gyc} path {N.wgrad_path(desc, N.PREC_BF16X3)}']
    for cl in (False, True):
        xs = x.contiguous(memory_format=torch.channels_last) if cl else x
        gs = gy.contiguous(memory_format=torch.channels_last) if cl else gy
        t = timeit(lambda: N.conv_wgrad(desc, xs, gs, N.PREC_BF16X3, gy_channels=gyc, channels_last=cl))
        tc = timeit(lambda: torch.ops.aten.convolution_backward(gs, xs, torch.empty(gyc, Cin, 3, 3, device=dev).contiguous(memory_format=torch.channels_last if cl else torch.contiguous_format),
                                                                None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [False, True, False]))
        gb = (x.numel() + gy.numel()) * 4 / 1e9
        row.append(f"{'nhwc' if cl else 'nchw'}: ours {t:.3f} ms ({gb / t * 1e3:.0f} GB/s of x+gy once) cudnn {tc:.3f} ms")
    print(' | '.join(row), flush=True)
