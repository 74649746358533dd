"""Print the tcgen05 planner's choice for a layer (host-only; works without a GPU).
usage: show_plan.py nd B Cin Cout D H W [k] [prec] [transposed]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200'))
from hebb import _native as N


def plan(nd, B, Cin, Cout, D, H, W, k=3, prec=1, tr=0):
    d = N.HebbDesc()
    d.nd, d.B, d.Cin, d.Cout, d.transposed = nd, B, Cin, Cout, tr
    dims = (D, H, W) if nd == 3 else (1, H, W)
    for i in range(3):
        on = i >= 3 - nd
        d.inp[i] = dims[i] if on else 1
        d.k[i] = k if on else 1
        d.stride[i] = (k if tr else 1) if on else 1
        d.pad_lo[i] = d.pad_hi[i] = (0 if tr else k // 2) if on else 0
    return N.plan(d, prec)


if __name__ == '__main__':
    a = list(map(int, sys.argv[1:]))
    print(plan(*a))
