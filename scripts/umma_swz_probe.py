"""Round-2 scouting (correctness, one CTA): MN-major operands in the SWIZZLE_64B / SWIZZLE_32B images --
[position][32 channels] (64 B per position) and [position][16 channels] (32 B per position) -- with position-shifted
starts and a ONE-position atom stride, i.e. the small-channel analogue of scripts/umma_sw128_probe.py: can one
instruction read several position-shifted copies of a 32- or 16-channel tile (taps as columns / rows)?"""
import ctypes, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200'))
import torch
from hebb import _native

LAYOUT = {128: 2, 64: 4, 32: 6}          # cute::UMMA::LayoutType
def desc_hi(lbo, sbo, layout): return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | (1 << 46) | (layout << 61)
def idesc(m, n, a_mn, b_mn): return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)
def bf16(a): return torch.from_numpy(a.astype(np.float32)).to(torch.bfloat16)
def u16(t): return t.view(torch.int16).numpy().view(np.uint16)

def image(T, W):
    """T: [positions][W/2 channels] bf16 -> byte image with 16-byte chunks XOR-ed with address bits 7.. (Swizzle<b,4,3>)"""
    P, C = T.shape
    nch = W // 16
    src = u16(T).reshape(P, nch, 8)
    img = np.zeros_like(src)
    for r in range(P):
        phase = ((r * W) >> 7) & (nch - 1)
        for c in range(nch):
            img[r, c ^ phase] = src[r, c]
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
    print(f'{name:86s} rel err {err:9.2e}  {"OK" if err < 1e-3 else "MISMATCH"}', flush=True)

rng = np.random.default_rng(0)
for W in (64, 32):
    C = W // 2                               # channels per atom
    lay = LAYOUT[W]
    P, KS = 64, 2
    K = 16 * KS
    X = bf16(rng.standard_normal((P, C))); R = bf16(rng.standard_normal((P, C)))
    Xf, Rf = X.float().numpy().astype(np.float64), R.float().numpy().astype(np.float64)
    a_img, b_img = image(X, W), image(R, W)
    sbo, kstep = 8 * W, 16 * W               # 8 positions per group, 16 positions per k-step
    nA = 128 // C                            # shifted copies of the x tile in the M rows
    for nB, sb in ((1, 0), (3, 0), (3, 5), (min(256 // C, 8), 1)):
        want = np.concatenate([Xf[j:j + K] for j in range(nA)], axis=1).T @ np.concatenate([Rf[sb + j:sb + j + K] for j in range(nB)], axis=1)
        got = run(a_img, b_img, desc_hi(W, sbo, lay), 0, kstep, desc_hi(W, sbo, lay), sb * W, kstep, 128, C * nB, KS)
        report(f'SWIZZLE_{W}B: M = {nA} copies x {C} ch (LBO = one position), N = {nB} copies x {C} ch, B start + {sb}', got, want)
    # copies one image row apart on the M side (kh taps), one position apart on the N side (kw taps)
    for d in (3, 10):
        want = np.concatenate([Xf[j * d:j * d + K] for j in range(min(nA, 3))] + [np.zeros((K, C))] * (nA - min(nA, 3)), axis=1)
        A = np.concatenate([Xf[j * d:j * d + K] for j in range(nA) if j * d + K <= P], axis=1)
        nv = A.shape[1]
        B = np.concatenate([Rf[j:j + K] for j in range(3)], axis=1)
        got = run(a_img, b_img, desc_hi(d * W, sbo, lay), 0, kstep, desc_hi(W, sbo, lay), 0, kstep, 128, C * 3, KS)
        report(f'SWIZZLE_{W}B: M copies {d} positions apart (first {nv // C} checked), N = 3 copies', got[:nv], A.T @ B)
