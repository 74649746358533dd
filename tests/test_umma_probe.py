"""Pins the shared-memory descriptor semantics the tensor-core kernels rely on: SWIZZLE_NONE
K-major and MN-major operands whose start address is shifted by whole 16-byte position slots, and MN-major
SWIZZLE_128B operands ([position][64 channels] images) with position-shifted starts and a one-position atom stride."""
import ctypes

import numpy as np
import pytest
import torch

from hebb import _native

pytestmark = pytest.mark.gpu


def desc_hi(lbo, sbo):
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | (1 << 46)


def idesc(m, n, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def bf16_round(a):
    t = torch.from_numpy(a.astype(np.float32)).to(torch.bfloat16)
    return t


def to_u16(t):
    return t.view(torch.int16).numpy().view(np.uint16)


def run_probe(a_img, b_img, a_hi, a_start, a_step, b_hi, b_start, b_step, idc, ksteps, m, n):
    lib = _native.load()
    a = torch.from_numpy(a_img.view(np.int16).copy()).cuda()
    b = torch.from_numpy(b_img.view(np.int16).copy()).cuda()
    out = torch.zeros(128, n, dtype=torch.float32, device='cuda')
    st = lib.hebb_debug_umma_probe(a.data_ptr(), a.numel() * 2, b.data_ptr(), b.numel() * 2,
                                   ctypes.c_uint64(a_hi), a_start, a_step, ctypes.c_uint64(b_hi), b_start, b_step,
                                   idc, ksteps, m, n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _native.check(st, 'probe')
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize('shift', [0, 1, 3, 8, 13])
def test_k_major_shifted_rows(shift):
    """Forward-kernel operand form: A[position, channel], rows 16 B apart, tap = start offset."""
    rng = np.random.default_rng(shift)
    M, N, K, rows = 128, 32, 32, 128 + 16
    A = bf16_round(rng.standard_normal((rows, K)))
    B = bf16_round(rng.standard_normal((N, K)))
    a_img = np.zeros((K // 8, rows, 8), np.uint16)      # [chunk][row][8]
    b_img = np.zeros((K // 8, N, 8), np.uint16)
    Au, Bu = to_u16(A), to_u16(B)
    for c in range(K // 8):
        a_img[c] = Au[:, c * 8:(c + 1) * 8]
        b_img[c] = Bu[:, c * 8:(c + 1) * 8]
    lbo_a, lbo_b = rows * 16, N * 16
    d = run_probe(a_img.reshape(-1), b_img.reshape(-1), desc_hi(lbo_a, 128), shift * 16, 2 * lbo_a,
                  desc_hi(lbo_b, 128), 0, 2 * lbo_b, idesc(M, N, 0, 0), K // 16, M, N)
    want = A.float().numpy()[shift:shift + M].astype(np.float64) @ B.float().numpy().astype(np.float64).T
    assert np.abs(d - want).max() < 1e-3 * np.abs(want).max()


@pytest.mark.parametrize('m', [64, 128])
@pytest.mark.parametrize('shift', [0, 1, 5, 8])
def test_mn_major_shifted_positions(m, shift):
    """dW-kernel operand form: A[channel, position] and B[channel, position], positions 16 B apart."""
    rng = np.random.default_rng(100 + shift + m)
    N, K, pos = 48, 64, 64 + 16
    A = bf16_round(rng.standard_normal((m, pos)))        # [ci][position]
    B = bf16_round(rng.standard_normal((N, K)))          # [co][position]
    Au, Bu = to_u16(A), to_u16(B)
    a_img = np.zeros((m // 8, pos, 8), np.uint16)        # [chunk][position][8 channels]
    b_img = np.zeros((N // 8, K, 8), np.uint16)
    for c in range(m // 8):
        a_img[c] = Au[c * 8:(c + 1) * 8, :].T
    for c in range(N // 8):
        b_img[c] = Bu[c * 8:(c + 1) * 8, :].T
    d = run_probe(a_img.reshape(-1), b_img.reshape(-1), desc_hi(128, pos * 16), shift * 16, 256,
                  desc_hi(128, K * 16), 0, 256, idesc(m, N, 1, 1), K // 16, m, N)
    want = A.float().numpy()[:, shift:shift + K].astype(np.float64) @ B.float().numpy().astype(np.float64).T
    if m == 128:
        got = d
    else:                                               # M=64: row r lives in lane (r%16) + 32*(r/16)
        lanes = [(r % 16) + 32 * (r // 16) for r in range(64)]
        got = d[lanes]
    assert np.abs(got - want).max() < 1e-3 * np.abs(want).max()


# ---- MN-major SWIZZLE_128B: the image is [position][64 channels], 16-byte chunks XOR-ed with (position % 8) ----
SW128 = 2


def desc_hi_sw(lbo, sbo):
    return desc_hi(lbo, sbo) | (SW128 << 61)


def sw128_image(T):
    """T: [positions][64] bf16 -> the swizzled shared-memory image as uint16."""
    P = T.shape[0]
    src = to_u16(T).reshape(P, 8, 8)
    img = np.zeros_like(src)
    for r in range(P):
        for c in range(8):
            img[r, c ^ (r & 7)] = src[r, c]
    return img.reshape(-1)


def _sw128_operands(seed, px=64, pr=64):
    rng = np.random.default_rng(seed)
    X0, X1 = bf16_round(rng.standard_normal((px, 64))), bf16_round(rng.standard_normal((px, 64)))
    R = bf16_round(rng.standard_normal((pr, 64)))
    a_img = np.concatenate([sw128_image(X0), sw128_image(X1)])
    f64 = lambda t: t.float().numpy().astype(np.float64)
    return a_img, sw128_image(R), f64(X0), f64(X1), f64(R), px * 128


@pytest.mark.parametrize('shift', [0, 1, 3, 8, 13])
def test_mn_major_sw128_shifted_start(shift):
    """The hardware swizzles on absolute shared-memory address bits: a start address moved by whole positions
    (128 B) reads the shifted tile with base_offset = 0 (the swizzled-response dW variant relies on it)."""
    a_img, b_img, X0, X1, R, atom_a = _sw128_operands(shift)
    K = 32
    d = run_probe(a_img, b_img, desc_hi_sw(atom_a, 1024), 0, 2048, desc_hi_sw(1024, 1024), shift * 128, 2048,
                  idesc(128, 64, 1, 1), K // 16, 128, 64)
    want = np.concatenate([X0[:K], X1[:K]], axis=1).T @ R[shift:shift + K]
    assert np.abs(d - want).max() < 1e-3 * np.abs(want).max()


@pytest.mark.parametrize('shift', [0, 2, 5])
def test_mn_major_sw128_one_position_atom_stride(shift):
    """LBO = 128 B: the 64-channel atoms of the B operand are ONE position apart, so an N = 192 instruction reads
    three position-shifted copies of the same response tile (the three taps of a kernel row) without replicas."""
    a_img, b_img, X0, X1, R, atom_a = _sw128_operands(100 + shift)
    K = 32
    d = run_probe(a_img, b_img, desc_hi_sw(atom_a, 1024), 0, 2048, desc_hi_sw(128, 1024), shift * 128, 2048,
                  idesc(128, 192, 1, 1), K // 16, 128, 192)
    B = np.concatenate([R[shift + j:shift + j + K] for j in range(3)], axis=1)
    want = np.concatenate([X0[:K], X1[:K]], axis=1).T @ B
    assert np.abs(d - want).max() < 1e-3 * np.abs(want).max()


@pytest.mark.parametrize('rows_apart', [1, 5, 10])
def test_mn_major_sw128_shifted_replica_rows(rows_apart):
    """The same on the A side: M = 128 = a 64-channel x tile and the tile `rows_apart` positions further (round-2
    layout: kh replicas without staging copies)."""
    a_img, b_img, X0, X1, R, atom_a = _sw128_operands(200 + rows_apart)
    K = 32
    d = run_probe(a_img, b_img, desc_hi_sw(rows_apart * 128, 1024), 0, 2048, desc_hi_sw(128, 1024), 0, 2048,
                  idesc(128, 192, 1, 1), K // 16, 128, 192)
    B = np.concatenate([R[j:j + K] for j in range(3)], axis=1)
    want = np.concatenate([X0[:K], X0[rows_apart:rows_apart + K]], axis=1).T @ B
    assert np.abs(d - want).max() < 1e-3 * np.abs(want).max()


# ---- SWIZZLE_64B / SWIZZLE_32B: [position][32 channels] and [position][16 channels] images ----
def swz_image(T, W):
    """T: [positions][W/2] bf16 -> image whose 16-byte chunks are XOR-ed with the address bits from bit 7 up."""
    P = T.shape[0]
    nch = W // 16
    src = to_u16(T).reshape(P, nch, 8)
    img = np.zeros_like(src)
    for r in range(P):
        phase = ((r * W) >> 7) & (nch - 1)
        for c in range(nch):
            img[r, c ^ phase] = src[r, c]
    return img.reshape(-1)


@pytest.mark.parametrize('W,layout', [(64, 4), (32, 6)])
@pytest.mark.parametrize('shift', [0, 5])
def test_mn_major_narrow_swizzles_one_position_atom_stride(W, layout, shift):
    """Small-channel analogue of the SWIZZLE_128B case (round-2 layout for the 16/32-channel layers): M = 128 rows =
    position-shifted copies of a 32- or 16-channel x tile, N = three position-shifted copies of the response tile."""
    C = W // 2
    rng = np.random.default_rng(W + shift)
    X, R = bf16_round(rng.standard_normal((64, C))), bf16_round(rng.standard_normal((64, C)))
    Xf, Rf = X.float().numpy().astype(np.float64), R.float().numpy().astype(np.float64)
    hi = lambda lbo: desc_hi(lbo, 8 * W) | (layout << 61)
    K, nA = 32, 128 // C
    d = run_probe(swz_image(X, W), swz_image(R, W), hi(W), 0, 16 * W, hi(W), shift * W, 16 * W,
                  idesc(128, 3 * C, 1, 1), K // 16, 128, 3 * C)
    A = np.concatenate([Xf[j:j + K] for j in range(nA)], axis=1)
    B = np.concatenate([Rf[shift + j:shift + j + K] for j in range(3)], axis=1)
    want = A.T @ B
    assert np.abs(d - want).max() < 1e-3 * np.abs(want).max()


# ---- K-major swizzled A operand over the SAME [position][channels] image (round-2 fused kernel: the forward reads
# the x tile as K-major rows, the update reads it as MN-major columns) ----
def kmajor_sw_image(T, W):
    """T: [rows][W/2] bf16 -> [rows][W bytes] image, 16-byte chunks XOR-ed with the address bits from bit 7 up."""
    return swz_image(T, W)


@pytest.mark.parametrize('W,layout', [(128, 2), (64, 4)])
@pytest.mark.parametrize('shift', [0, 1, 3, 8, 13])
@pytest.mark.parametrize('lbo', [0, 16])
def test_k_major_swizzled_shifted_rows_and_k_offsets(W, layout, shift, lbo):
    """A = 128 positions x 16 channels read K-major out of a swizzled [position][W bytes] image: the start address is
    moved by whole rows (a tap) and by 32 bytes inside the row (the next 16 channels / the lo half), base_offset 0."""
    rng = np.random.default_rng(W + shift)
    C = W // 2                                    # bf16 values per row
    rows, M, N = 128 + 16, 128, 32
    A = bf16_round(rng.standard_normal((rows, C)))
    ksteps = C // 16
    B = bf16_round(rng.standard_normal((N, C)))
    Bu = to_u16(B)
    b_img = np.zeros((C // 8, N, 8), np.uint16)
    for c in range(C // 8):
        b_img[c] = Bu[:, c * 8:(c + 1) * 8]
    lbo_b = N * 16
    a_hi = desc_hi(lbo, 8 * W) | (layout << 61)
    d = run_probe(kmajor_sw_image(A, W), b_img.reshape(-1), a_hi, shift * W, 32, desc_hi(lbo_b, 128), 0, 2 * lbo_b,
                  idesc(M, N, 0, 0), ksteps, M, N)
    want = A.float().numpy()[shift:shift + M].astype(np.float64) @ B.float().numpy().astype(np.float64).T
    assert np.abs(d - want).max() < 1e-3 * np.abs(want).max()


@pytest.mark.parametrize('W,layout', [(128, 2), (64, 4)])
def test_k_major_swizzled_single_k_slice(W, layout):
    """One K = 16 slice in the middle of the row (e.g. only the lo half): start + 32 * j, one instruction."""
    rng = np.random.default_rng(W)
    C = W // 2
    rows, M, N = 128 + 8, 128, 16
    A = bf16_round(rng.standard_normal((rows, C)))
    B = bf16_round(rng.standard_normal((N, 16)))
    Bu = to_u16(B)
    b_img = np.zeros((2, N, 8), np.uint16)
    for c in range(2):
        b_img[c] = Bu[:, c * 8:(c + 1) * 8]
    a_hi = desc_hi(0, 8 * W) | (layout << 61)
    for j in range(C // 16):
        d = run_probe(kmajor_sw_image(A, W), b_img.reshape(-1), a_hi, 5 * W + 32 * j, 0, desc_hi(N * 16, 128), 0, 0,
                      idesc(M, N, 0, 0), 1, M, N)
        want = A.float().numpy()[5:5 + M, 16 * j:16 * j + 16].astype(np.float64) @ B.float().numpy().astype(np.float64).T
        assert np.abs(d - want).max() < 1e-3 * np.abs(want).max(), j
