"""Round-2 scouting: cycles per tcgen05.mma with SWIZZLE_128B operands vs the SWIZZLE_NONE forms used now
(timing only, zero data, 64 MMAs per commit so the barrier round trip is amortised)."""
import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200'))
import torch
from hebb import _native

def desc_hi(lbo, sbo, layout=0): return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | (1 << 46) | (layout << 61)
def idesc(m, n, a_mn, b_mn): return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)
SW128 = 2
lib = _native.load()
REGION = 98304
def run(name, a_hi, a_step, b_hi, b_step, m, n, a_mn, b_mn, per_round=64, iters=40, ctas=148):
    cyc = torch.zeros(ctas, dtype=torch.int64, device='cuda')
    st = lib.hebb_debug_umma_rate(ctypes.c_uint64(a_hi), a_step, ctypes.c_uint64(b_hi), b_step, REGION, idesc(m, n, a_mn, b_mn),
                                  per_round, iters, n, ctas, cyc.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _native.check(st, name); torch.cuda.synchronize()
    c = cyc.float().mean().item() / (per_round * iters)
    print(f'{name:52s} M={m:3d} N={n:3d}: {c:7.1f} cyc/MMA  (math floor {max(m,128)*n/256:5.1f})', flush=True)

for n in (16, 32, 64, 128, 256):
    # a_step cycles through the 4 K-steps of a 64-wide swizzle atom and 8 row groups (stays inside REGION)
    run('K-major NONE  A(LBO=4160,SBO=128) B(LBO=n*16)', desc_hi(4160, 128), 16, desc_hi(n * 16, 128), 0, 128, n, 0, 0)
    run('K-major SW128 A,B rows of 128 B (SBO=1024)', desc_hi(0, 1024, SW128), 32, desc_hi(0, 1024, SW128), 32, 128, n, 0, 0, per_round=4, iters=640)
    run('K-major SW128, same start every MMA', desc_hi(0, 1024, SW128), 0, desc_hi(0, 1024, SW128), 0, 128, n, 0, 0)
    for m in (64, 128):
        run('MN-major NONE A(SBO=4160) B(SBO=2048)', desc_hi(128, 4160), 256, desc_hi(128, 2048), 256, m, n, 1, 1, per_round=16, iters=160)
        run('MN-major SW128 (LBO=16384,SBO=1024)', desc_hi(16384, 1024, SW128), 2048, desc_hi(16384, 1024, SW128), 2048, m, n, 1, 1, per_round=8, iters=320)
        run('MN-major SW128, same start every MMA', desc_hi(16384, 1024, SW128), 0, desc_hi(16384, 1024, SW128), 0, m, n, 1, 1)
