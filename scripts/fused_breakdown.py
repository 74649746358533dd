"""Times one fused-kernel layer (HEBB_FUSED_DBG is read once per process: run once per mask)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200')); sys.path.insert(0, ROOT)
import torch
import hebb
from hebb import _native
Cin, Cout, H, k = [int(v) for v in sys.argv[1:5]]
UPD = 0 if (len(sys.argv) > 5 and sys.argv[5] == 'fwd') else 1      # 'fwd': forward only (no plasticity update)
B = 64
x = torch.randn(B, Cin, H, H, device='cuda')
layer = hebb.HebbianConv2d(Cin, Cout, k, padding=k // 2, bias=False, k=50., alpha=1.).cuda().train()
layer.prec = 'bf16x3'
desc = layer._desc(x.shape, True)
w = layer.weight.detach(); dw = torch.zeros_like(w); y = torch.empty(B, Cout, H, H, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
ts = []
for i in range(6):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _native.conv_step(desc, x, w, None, 50., y, None, dw, _native.F_WNRM | (_native.F_UPDATE if UPD else 0), 1)
    e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"dbg={os.environ.get('HEBB_FUSED_DBG', '0'):>3s} {'upd' if UPD else 'fwd'} fused={os.environ.get('HEBB_FUSED', '1')} {Cin}->{Cout} k{k} @{H}: {min(ts[1:]):.3f} ms  plan {_native.fused_plan(desc)}", flush=True)

if os.environ.get('HEBB_FUSED_PROF') == '1':
    import ctypes
    import numpy as np
    grid = _native.fused_plan(desc)['grid']
    buf = (ctypes.c_longlong * (grid * 165))()
    lib = _native.load()
    lib.hebb_debug_fused_prof.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
    n = lib.hebb_debug_fused_prof(buf, grid * 165)
    t = np.array(buf[:n], dtype=np.float64).reshape(grid, 15, 11)
    names = {0: 'w_full', 1: 'st_empty', 2: 'xr_full', 3: 'tf_empty', 4: 'r_full', 5: 'st_full', 6: 'xr_empty', 7: 'tf_full', 8: 'done', 9: 'r_empty'}
    roles = {0: 'producer', 1: 'mma.fwd', 2: 'mma.dw', 3: 'conv0', 6: 'conv3', 7: 'epi0.a', 11: 'epi1.a'}
    m = t.mean(axis=0)
    if Cin <= 3:
        roles = {0: 'producer', 1: 'mma.fwd', 2: 'mma.dw', 3: 'conv0', 6: 'conv3', 7: 'epi0.a', 11: 'epi1.a'}
    else:
        roles = {0: 'producer', 1: 'mma.fwd', 2: 'mma.dw', 3: 'conv0', 4: 'conv1', 5: 'epi0.a', 9: 'epi1.a'}
    ep = 7 if Cin <= 3 else 5
    print('   epilogue sections (kcyc): tmem-load %.1f  y/winner/sums %.1f  softmax/split/stores %.1f  fence+arrive %.1f' % tuple(m[ep, i] / 1e3 for i in range(4)))
    for w, rn in roles.items():
        tot = m[w, 10]
        print(f'   {rn:9s} total {tot / 1e3:8.1f} kcyc | ' + '  '.join(f'{names[i]} {m[w, i] / 1e3:.1f}' for i in range(10) if m[w, i] > 0.005 * tot))
