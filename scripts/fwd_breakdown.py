"""Forward-kernel time breakdown with the HEBB_FWD_DBG knob (set in the environment by the caller).
usage: HEBB_FWD_DBG=n python scripts/fwd_breakdown.py   -> prints fwd-only stage time for a few 3-D layers"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200'))
import torch, hebb
from hebb import _native

LAYERS = [  # (cls, Cin, Cout, k, B, spatial)
    ('c3', 1, 64, 3, 4, (96, 96, 80)), ('c3', 64, 64, 3, 4, (96, 96, 80)), ('c3', 128, 64, 3, 4, (96, 96, 80)),
    ('c3', 128, 128, 3, 8, (48, 48, 40)), ('c3', 256, 256, 3, 8, (24, 24, 20)), ('t3', 128, 64, 2, 8, (48, 48, 40)),
    ('t3', 256, 128, 2, 8, (24, 24, 20)), ('c3', 1024, 1024, 3, 8, (6, 6, 5)), ('c3', 512, 512, 3, 8, (12, 12, 10)),
    ('c3', 256, 512, 3, 8, (12, 12, 10)), ('c3', 1024, 512, 3, 8, (12, 12, 10)),
    ('c2', 16, 16, 3, 64, (256, 256)), ('c2', 32, 16, 3, 64, (256, 256)), ('c2', 64, 64, 3, 64, (64, 64)),
]
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
out = []
for kind, ci, co, k, B, sp in LAYERS:
    if kind == 'c3':
        m = hebb.HebbianConv3d(ci, co, k, padding=k // 2, bias=False, k=50., alpha=1.)
    elif kind == 't3':
        m = hebb.HebbianConvTranspose3d(ci, co, k, stride=k, bias=False, k=50., alpha=1.)
    else:
        m = hebb.HebbianConv2d(ci, co, k, padding=k // 2, bias=False, k=50., alpha=1.)
    m = m.cuda().train()
    x = torch.randn(B, ci, *sp, device='cuda')
    with torch.no_grad():
        y = m(x)
    desc = m._desc(x.shape, True)
    prec = _native.parse_prec(m.prec)
    w = m._raw(m.weight.detach())
    dw = torch.zeros_like(w)
    yb = torch.empty_like(y)
    base = _native.F_WNRM | _native.F_UPDATE
    res = {}
    for name, fl in (('fwd', _native.F_ONLY_FWD), ('dw', _native.F_ONLY_DW)):
        best = 1e9
        for rep in range(3):
            flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _native.conv_step(desc, x, w, m.bias.detach(), float(m.k), yb, None, dw, base | fl, prec)
            e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[name] = round(best, 3)
    print('dbg', os.environ.get('HEBB_FWD_DBG', '0'), kind, ci, co, B, sp, res, flush=True)
    del m, x, y, yb, dw
