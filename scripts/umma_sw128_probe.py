"""Round-2 scouting (correctness, one CTA): MN-major SWIZZLE_128B operands whose image is [position][64 channels]
(128 bytes per position, 16-byte chunks XOR-ed with position % 8).  Questions:
  1. does the canonical form (LBO = stride between 64-channel atoms, SBO = 1024 B between 8-position groups) work;
  2. can the start address be shifted by whole positions (128 B), with base_offset 0 or (shift % 8);
  3. can LBO be ONE position (128 B), so that an N = 192 instruction reads three position-shifted copies of the
     same 64-channel response tile (taps kw = 0, 1, 2 without replicas), and likewise M = 128 = two shifted copies
     of a 64-channel x tile."""
import ctypes, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200'))
import torch
from hebb import _native

SW128 = 2
def desc_hi(lbo, sbo, layout=SW128, base_offset=0):
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | (1 << 46) | (base_offset & 7) << 49 | (layout << 61)
def idesc(m, n, a_mn, b_mn): return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)
def bf16(a): return torch.from_numpy(a.astype(np.float32)).to(torch.bfloat16)
def u16(t): return t.view(torch.int16).numpy().view(np.uint16)

def image(T):
    """T: [positions][64] bf16 -> swizzled byte image [positions][8 chunks][8]"""
    P = T.shape[0]
    src = u16(T).reshape(P, 8, 8)
    img = np.zeros_like(src)
    for r in range(P):
        for c in range(8):
            img[r, c ^ (r & 7)] = src[r, c]
    return img.reshape(-1)

def run(a_img, b_img, a_hi, a_start, a_step, b_hi, b_start, b_step, m, n, ksteps):
    lib = _native.load()
    a = torch.from_numpy(a_img.view(np.int16).copy()).cuda(); b = torch.from_numpy(b_img.view(np.int16).copy()).cuda()
    out = torch.zeros(128, n, dtype=torch.float32, device='cuda')
    st = lib.hebb_debug_umma_probe(a.data_ptr(), a.numel() * 2, b.data_ptr(), b.numel() * 2, ctypes.c_uint64(a_hi), a_start, a_step,
                                   ctypes.c_uint64(b_hi), b_start, b_step, idesc(m, n, 1, 1), ksteps, m, n, out.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream)
    _native.check(st, 'probe'); torch.cuda.synchronize()
    return out.cpu().numpy()

def report(name, got, want):
    err = np.abs(got - want).max() / np.abs(want).max()
    print(f'{name:70s} rel err {err:9.2e}  {"OK" if err < 1e-3 else "MISMATCH"}', flush=True)

rng = np.random.default_rng(0)
PX, PR, KS = 64, 64, 2                    # positions held per image; 2 k-steps of 16 positions
X0 = bf16(rng.standard_normal((PX, 64))); X1 = bf16(rng.standard_normal((PX, 64)))      # channels 0-63, 64-127
R = bf16(rng.standard_normal((PR, 64)))
Xf = [X0.float().numpy().astype(np.float64), X1.float().numpy().astype(np.float64)]
Rf = R.float().numpy().astype(np.float64)
a_img = np.concatenate([image(X0), image(X1)]); atomA = PX * 128
b_img = image(R)
K = 16 * KS
def want(sa_, sb_, nrep=1, arep=None):
    rows = []
    if arep is None: rows = [Xf[0][sa_:sa_ + K], Xf[1][sa_:sa_ + K]]
    else: rows = [Xf[0][sa_:sa_ + K], Xf[0][sa_ + arep:sa_ + arep + K]]
    A = np.concatenate(rows, axis=1)                       # [K][128]
    B = np.concatenate([Rf[sb_ + j:sb_ + j + K] for j in range(nrep)], axis=1)   # [K][64*nrep]
    return A.T @ B

# 1. canonical
report('1. canonical: M=128 (2 atoms), N=64', run(a_img, b_img, desc_hi(atomA, 1024), 0, 2048, desc_hi(1024, 1024), 0, 2048, 128, 64, KS), want(0, 0))
# 2. shifted starts
for s in (1, 3, 8, 13):
    for bo in (0, s & 7):
        report(f'2. B start + {s} positions, base_offset {bo}', run(a_img, b_img, desc_hi(atomA, 1024), 0, 2048, desc_hi(1024, 1024, base_offset=bo), s * 128, 2048, 128, 64, KS), want(0, s))
        report(f'2. A start + {s} positions, base_offset {bo}', run(a_img, b_img, desc_hi(atomA, 1024, base_offset=bo), s * 128, 2048, desc_hi(1024, 1024), 0, 2048, 128, 64, KS), want(s, 0))
# 3. one-position atom stride: N = 192 = three shifted copies of the response tile
for s in (0, 2, 5):
    report(f'3. N=192, LBO = 128 B, B start + {s}', run(a_img, b_img, desc_hi(atomA, 1024), 0, 2048, desc_hi(128, 1024), s * 128, 2048, 128, 192, KS), want(0, s, nrep=3))
    report(f'3. N=128, LBO = 128 B, B start + {s}', run(a_img, b_img, desc_hi(atomA, 1024), 0, 2048, desc_hi(128, 1024), s * 128, 2048, 128, 128, KS), want(0, s, nrep=2))
# 4. M = 128 = the 64-channel x tile and a copy shifted by `d` positions (a kernel row apart)
for d in (1, 5, 10):
    report(f'4. M=128 = x and x shifted by {d} (LBO = {d}*128 B), N=192', run(a_img, b_img, desc_hi(d * 128, 1024), 0, 2048, desc_hi(128, 1024), 0, 2048, 128, 192, KS), want(0, 0, nrep=3, arep=d))
