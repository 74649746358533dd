"""Cycles per tcgen05.mma (bf16, fp32 accumulate) for operand layouts of interest (timing only; zeros)."""
import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200'))
import torch
from hebb import _native

def desc_hi(lbo, sbo, layout=0): return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | (1 << 46) | (layout << 61)
def idesc(m, n, a_mn, b_mn): return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)
SW128, SW64, SW32 = 2, 4, 6
lib = _native.load()
REGION = 98304
def run(name, a_hi, a_step, b_hi, b_step, m, n, a_mn, b_mn, per_round=16, iters=100, ctas=1):
    cyc = torch.zeros(ctas, dtype=torch.int64, device='cuda')
    st = lib.hebb_debug_umma_rate(ctypes.c_uint64(a_hi), a_step, ctypes.c_uint64(b_hi), b_step, REGION, idesc(m, n, a_mn, b_mn),
                                  per_round, iters, n, ctas, cyc.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _native.check(st, name); torch.cuda.synchronize()
    c = cyc.float().mean().item() / (per_round * iters)
    print(f'{name:44s} M={m:3d} N={n:3d} ctas={ctas:3d}: {c:7.1f} cyc/MMA  (math floor {max(m,128)*n/256:5.1f})', flush=True)

for ctas in (1, 148):
    for n in (16, 64, 128, 256):
        run('K-major none  A(LBO=4160,SBO=128)', desc_hi(4160, 128), 16, desc_hi(n * 16, 128), 0, 128, n, 0, 0, ctas=ctas)
    for n in (16, 64, 128, 256):
        for m in (64, 128):
            # A: m/8 chunks at 4160 B ; B: n/8 chunks at 2048 B  (max extent 32*2048 = 64 KB < REGION)
            run('MN-major none A(SBO=4160) B(SBO=2048)', desc_hi(128, 4160), 256, desc_hi(128, 2048), 256, m, n, 1, 1, ctas=ctas)
    for n in (16, 64, 128, 256):
        # SW128 K-major: rows of 128 B (64 bf16 of K), 8-row groups 1024 B apart; K advance 32 B inside the atom
        run('K-major SW128 (SBO=1024)', desc_hi(0, 1024, SW128), 32, desc_hi(0, 1024, SW128), 32, 128, n, 0, 0, per_round=4, iters=400, ctas=ctas)
    for n in (16, 64, 128, 256):
        for m in (64, 128):
            # SW128 MN-major: rows of 128 B = 64 channels of one position; K = 16 positions = 2 groups of 8 rows (SBO=1024);
            # 64-channel slabs LBO apart (A: 16 KB, B: 16 KB)
            run('MN-major SW128 (LBO=16384,SBO=1024)', desc_hi(16384, 1024, SW128), 2048, desc_hi(16384, 1024, SW128), 2048, m, n, 1, 1, per_round=8, iters=200, ctas=ctas)
    for n in (16, 32):
        run('MN-major SW32 (LBO=8192,SBO=256) Cin16', desc_hi(8192, 256, SW32), 512, desc_hi(8192, 256, SW32), 512, 64, n, 1, 1, ctas=ctas)
        run('K-major SW32 (SBO=256) K=16 rows 32B', desc_hi(0, 256, SW32), 4096, desc_hi(0, 256, SW32), 0, 128, n, 0, 0, per_round=8, iters=200, ctas=ctas)
