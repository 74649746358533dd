"""Wait-cycle table of the fused kernel in weight-gradient mode (HEBB_FUSED_PROF=1; Cin = 32 layers, 4 converter warps).
usage: HEBB_FUSED_PROF=1 wgrad_breakdown.py Cout gy_channels nchw|nhwc"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200'))
import numpy as np
import torch
from hebb import _native as N

Cout, gyc, fmt = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
B, Cin = 64, 32
dev = torch.device('cuda', 0)
x = torch.randn(B, Cin, 256, 256, device=dev)
gy = torch.randn(B, gyc, 256, 256, device=dev)
cl = fmt == 'nhwc'
if cl:
    x, gy = x.contiguous(memory_format=torch.channels_last), gy.contiguous(memory_format=torch.channels_last)
desc = N.make_desc(2, B, Cin, Cout, (256, 256), (3, 3), (1, 1), (1, 1), (1, 1), False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(4):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); N.conv_wgrad(desc, x, gy, N.PREC_BF16X3, gy_channels=gyc, channels_last=cl); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
sub = N.make_desc(2, B, 32, 16 if Cout == 16 else 32, (256, 256), (3, 3), (1, 1), (1, 1), (1, 1), False)
plan = N.fused_plan(sub)
print(f'wgrad 32->{gyc} (padded {Cout}) {fmt}: {min(ts[1:]):.3f} ms  plan {plan}', flush=True)
if os.environ.get('HEBB_FUSED_PROF') == '1':
    grid = plan['grid']
    buf = (ctypes.c_longlong * (grid * 165))()
    lib = N.load()
    lib.hebb_debug_fused_prof.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
    n = lib.hebb_debug_fused_prof(buf, grid * 165)
    t = np.array(buf[:n], dtype=np.float64).reshape(grid, 15, 11)
    m = t.mean(axis=0)
    names = {0: 'w_full', 1: 'st_empty', 2: 'xr_full', 3: 'tf_empty', 4: 'r_full', 5: 'st_full', 6: 'xr_empty', 7: 'tf_full', 8: 'done', 9: 'r_empty'}
    roles = {0: 'producer', 2: 'mma.dw', 3: 'conv0', 6: 'conv3', 7: 'epi0.a', 11: 'epi1.a'}
    print('   gy-loader sections (kcyc): load+split+stores %.1f  fence+arrive %.1f' % (m[7, 2] / 1e3, m[7, 3] / 1e3))
    print('   converter sections (kcyc, conv0): conversion loop %.1f  fence+arrives %.1f' % (m[3, 0] / 1e3, m[3, 1] / 1e3))
    for w, rn in roles.items():
        tot = m[w, 10]
        skip = (0, 1, 2, 3) if w >= 7 else ((0, 1) if w >= 3 else ())
        print(f'   {rn:9s} total {tot / 1e3:8.1f} kcyc | ' + '  '.join(f'{names[i]} {m[w, i] / 1e3:.1f}' for i in range(10) if i not in skip and m[w, i] > 0.005 * tot))
