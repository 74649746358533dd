"""One small layer through the fused kernel, compared with the CPU oracle; prints the watchdog code on failure."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200')); sys.path.insert(0, ROOT)
import torch
import hebb
from hebb import _native
from oracle import hebb_oracle as O

cases = [(4, 3, 16, 3, 1, (40, 36), 50.0, True), (2, 1, 32, 3, 1, (21, 44), 20.0, True), (2, 3, 16, 3, 1, (64, 256), 50.0, True), (2, 16, 16, 3, 0, (20, 24), 5.0, False), (2, 16, 16, 3, 0, (20, 24), 5.0, True), (2, 16, 16, 3, 1, (20, 24), 5.0, True), (2, 16, 16, 3, 1, (20, 24), 5.0, False), (2, 32, 32, 3, 1, (20, 24), 5.0, True),
         (3, 16, 32, 3, 1, (40, 36), 50.0, True), (2, 32, 16, 1, 0, (24, 28), 10.0, True), (1, 16, 16, 3, 1, (19, 256), 50.0, True)]
only = int(sys.argv[1]) if len(sys.argv) > 1 else -1
for i, (B, Cin, Cout, k, pad, sp, kinv, upd) in enumerate(cases):
    if only >= 0 and i != only:
        continue
    g = torch.Generator().manual_seed(i)
    x = torch.randn(B, Cin, *sp, generator=g)
    layer = hebb.HebbianConv2d(Cin, Cout, k, padding=pad, bias=True, k=kinv, alpha=1.)
    with torch.no_grad():
        layer.bias.copy_(torch.randn(Cout, generator=g) * 0.05)
    w, b = layer.weight.detach().clone(), layer.bias.detach().clone()
    xp = O.zero_halo(x, pad, 2)
    y_ref = O.conv_activation(xp, w, b, (1, 1))
    dw_ref = O.swta_delta(xp, y_ref, w, kinv, (1, 1))
    layer.prec = 'bf16x3'
    layer.record_winners = True
    layer = layer.cuda()
    layer.train(upd)
    d = layer._desc(x.shape, True)
    print('case', i, (B, Cin, Cout, k, pad, sp, upd), 'path', _native.layer_path(d, 1, 3), 'plan', _native.fused_plan(d), flush=True)
    try:
        y = layer(x.cuda())
        torch.cuda.synchronize()
    except Exception as e:
        print('FAILED:', str(e)[:300], 'watchdog', _native.load().hebb_watchdog_code(), flush=True)
        sys.exit(1)
    ey = float((y.cpu() - y_ref).norm() / y_ref.norm())
    ed = float((layer.delta_w.cpu() - dw_ref).norm() / dw_ref.norm()) if upd else 0.0
    wm = int((layer.winners.cpu().long() != y_ref.argmax(1)).sum())
    print(f'   y err {ey:.2e}  dW err {ed:.2e}  winner mismatches {wm}', flush=True)
    if ey > 1e-4:
        err = (y.cpu() - y_ref).abs().amax(dim=(0, 1))
        print('   per-pixel max err map (rows with err > 1e-3):', [(r, [c for c in range(err.shape[1]) if err[r, c] > 1e-3][:8]) for r in range(err.shape[0]) if (err[r] > 1e-3).any()][:12])
