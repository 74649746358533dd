"""Cycles per tcgen05.mma when consecutive instructions share their A tile (MN-major SWIZZLE_NONE operands as in the
dW kernel): groups of G MMAs with one A descriptor and G different B descriptors, with and without the A-operand
collector hints.  Timing only (zero data), 148 CTAs."""
import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200'))
import torch
from hebb import _native

def desc_hi(lbo, sbo, layout=0): return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | (1 << 46) | (layout << 61)
def idesc(m, n, a_mn, b_mn): return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)
lib = _native.load()
REGION = 98304
def run(m, n, group, hint, b_step=16, d_step=0, a_sbo=4160, b_sbo=2304, per_round=72, iters=40, ctas=148):
    cyc = torch.zeros(ctas, dtype=torch.int64, device='cuda')
    # A: M/8 chunks a_sbo B apart, 16 positions = 256 B per k-step; B: N/8 chunks b_sbo B apart, taps shift by b_step B
    st = lib.hebb_debug_umma_rate_shared_a(ctypes.c_uint64(desc_hi(128, a_sbo)), 256, ctypes.c_uint64(desc_hi(128, b_sbo)), b_step,
                                           REGION, idesc(m, n, 1, 1), per_round, group, hint, d_step, iters, n, ctas,
                                           cyc.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _native.check(st, 'rate'); torch.cuda.synchronize()
    c = cyc.float().mean().item() / (per_round * iters)
    print(f'M={m:3d} N={n:3d} group={group:2d} hints={hint} b_step={b_step:3d} d_step={d_step:3d} sbo={a_sbo}/{b_sbo}: {c:6.1f} cyc/MMA, {c * group:7.1f} per group', flush=True)

for m, n in ((128, 64), (128, 32), (64, 64), (128, 16)):
    for group in (2, 3, 4, 6, 8, 9, 18):
        if n * group > 512 and group != 18: continue
        for hint in (0, 1, 2):
            run(m, n, group, hint, d_step=n if n * group <= 512 else 0)
    for hint in (0, 1):
        run(m, n, 6, hint, b_step=256, b_sbo=4160)       # aligned B starts
        run(m, n, 6, hint, b_step=128, b_sbo=4160)
        run(m, n, 6, hint, b_step=16, d_step=0)           # same accumulator
