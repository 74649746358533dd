"""ctypes binding of libhebb_sm100.so (C ABI in include/hebb_sm100.h).

PyTorch is used here only for device memory and streams.  There is no CPU path:
if the shared library cannot be loaded, or no sm_100 device is present, every
compute entry raises RuntimeError.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libhebb_sm100.so')

PREC_FP32, PREC_BF16X3, PREC_BF16 = 0, 1, 2
_PREC_NAMES = {'fp32': PREC_FP32, 'bf16x3': PREC_BF16X3, 'fp32x3': PREC_BF16X3, 'bf16': PREC_BF16}
F_UPDATE, F_WNRM, F_RULE_HPCA = 1, 2, 4
F_ONLY_PACK, F_ONLY_FWD, F_ONLY_DW = 0x100, 0x200, 0x400

EXPORTS = [
    'hebb_query', 'hebb_status_str', 'hebb_last_cuda_error', 'hebb_version', 'hebb_out_shape',
    'hebb_workspace_bytes', 'hebb_wnorm', 'hebb_conv_swta_step', 'hebb_convT_swta_step',
    'hebb_local_update_multi', 'hebb_debug_umma_probe', 'hebb_debug_launch_count', 'hebb_uses_tensor_cores',
    'hebb_debug_umma_rate', 'hebb_debug_umma_rate_shared_a', 'hebb_debug_plan', 'hebb_bn_act_train', 'hebb_bn_act_train_dropout', 'hebb_bn_act_from_stats_dropout', 'hebb_upsample2x_bilinear',
    'hebb_layer_path', 'hebb_wgrad_path', 'hebb_debug_fused_plan', 'hebb_watchdog_code', 'hebb_debug_fused_prof', 'hebb_conv_swta_step_stats', 'hebb_bn_act_from_stats', 'hebb_conv_wgrad', 'hebb_maxpool2x',
    'hebb_bias_relu_dropout', 'hebb_bias_relu_dropout_state', 'hebb_mask_scale', 'hebb_mask_scale_gb',
]


class HebbDesc(ctypes.Structure):
    _fields_ = [('nd', ctypes.c_int32), ('B', ctypes.c_int32), ('Cin', ctypes.c_int32), ('Cout', ctypes.c_int32),
                ('inp', ctypes.c_int32 * 3), ('k', ctypes.c_int32 * 3), ('stride', ctypes.c_int32 * 3),
                ('pad_lo', ctypes.c_int32 * 3), ('pad_hi', ctypes.c_int32 * 3), ('transposed', ctypes.c_int32)]


_lib = None
_lib_lock = threading.Lock()
_default_prec = _PREC_NAMES.get(os.environ.get('HEBB_PREC', 'bf16x3').lower(), PREC_BF16X3)


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load the shared library (once).  Raises RuntimeError if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f'{_LIB_PATH} not found: build it with `python build.py` (or __graft_entry__.build()). '
                'The Hebbian layers have no CPU or PyTorch fallback.')
        lib = ctypes.CDLL(_LIB_PATH)
        vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
        lib.hebb_query.argtypes = [ctypes.POINTER(i32)] * 3
        lib.hebb_status_str.restype = ctypes.c_char_p
        lib.hebb_status_str.argtypes = [i32]
        lib.hebb_version.restype = ctypes.c_char_p
        lib.hebb_out_shape.argtypes = [ctypes.POINTER(HebbDesc), ctypes.POINTER(ctypes.c_int32 * 3)]
        lib.hebb_workspace_bytes.argtypes = [ctypes.POINTER(HebbDesc), i32, ctypes.POINTER(ctypes.c_size_t)]
        lib.hebb_wnorm.argtypes = [vp, vp, vp, i64, i64, i64, i64, i64, vp]
        step = [ctypes.POINTER(HebbDesc), vp, vp, vp, f32, vp, vp, vp, vp, ctypes.c_size_t, ctypes.c_uint, i32, vp]
        lib.hebb_conv_swta_step.argtypes = step
        lib.hebb_convT_swta_step.argtypes = step
        lib.hebb_conv_swta_step_stats.argtypes = step[:-1] + [vp, ctypes.POINTER(i32), vp]
        lib.hebb_bn_act_from_stats.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, f32, f32, f32, vp, ctypes.c_size_t, vp]
        lib.hebb_bn_act_from_stats_dropout.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, f32, f32, f32, f32, vp, vp, ctypes.c_size_t, vp]
        lib.hebb_conv_wgrad.argtypes = [ctypes.POINTER(HebbDesc), vp, vp, vp, i32, i32, vp, ctypes.c_size_t, i32, vp]
        lib.hebb_local_update_multi.argtypes = [i32, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(i64),
                                                ctypes.POINTER(f32), ctypes.POINTER(ctypes.c_int32), vp]
        lib.hebb_debug_umma_probe.argtypes = [vp, i32, vp, i32, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32,
                                              ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                              i32, i32, i32, vp, vp]
        lib.hebb_debug_umma_rate.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32,
                                             ctypes.c_uint32, ctypes.c_uint32, i32, i32, i32, i32, vp, vp]
        lib.hebb_debug_umma_rate_shared_a.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32,
                                                      ctypes.c_uint32, ctypes.c_uint32, i32, i32, i32, i32, i32, i32, i32, vp, vp]
        lib.hebb_bn_act_train.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, f32, f32, f32, vp, ctypes.c_size_t, vp]
        lib.hebb_bn_act_train_dropout.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, f32, f32, f32, f32, vp, vp, ctypes.c_size_t, vp]
        lib.hebb_upsample2x_bilinear.argtypes = [vp, vp, i64, i64, i64, vp]
        lib.hebb_maxpool2x.argtypes = [vp, vp, i64, i64, i64, i64, i32, vp]
        lib.hebb_bias_relu_dropout.argtypes = [vp, vp, vp, vp, i64, i64, i64, f32, ctypes.c_uint64, vp]
        lib.hebb_bias_relu_dropout_state.argtypes = [vp, vp, vp, vp, i64, i64, i64, f32, vp, vp]
        lib.hebb_mask_scale_gb.argtypes = [vp, vp, vp, vp, i64, i64, f32, vp, i64, vp]
        lib.hebb_mask_scale.argtypes = [vp, vp, vp, i64, f32, vp]
        lib.hebb_debug_launch_count.restype = ctypes.c_ulonglong
        lib.hebb_uses_tensor_cores.argtypes = [ctypes.POINTER(HebbDesc), i32]
        lib.hebb_layer_path.argtypes = [ctypes.POINTER(HebbDesc), i32, ctypes.c_uint]
        lib.hebb_wgrad_path.argtypes = [ctypes.POINTER(HebbDesc), i32]
        lib.hebb_debug_fused_plan.argtypes = [ctypes.POINTER(HebbDesc), ctypes.POINTER(ctypes.c_int), i32]
        for name in EXPORTS:
            getattr(lib, name)      # fail loudly if a declared symbol is missing
        _lib = lib
    return _lib


def check(status: int, what: str = ''):
    if status != 0:
        lib = load()
        msg = lib.hebb_status_str(status).decode()
        extra = ''
        if status == -5:
            extra = f' (cudaError {lib.hebb_last_cuda_error()})'
            wd = lib.hebb_watchdog_code()
            if wd:                       # a kernel's bounded wait timed out and trapped: HEBB_EKERNEL
                msg = lib.hebb_status_str(-7).decode()
                extra += f' (watchdog code {wd})'
        raise RuntimeError(f'libhebb_sm100: {what}: {msg}{extra}')


def parse_prec(p) -> int:
    if p is None:
        return _default_prec
    if isinstance(p, str):
        try:
            return _PREC_NAMES[p.lower()]
        except KeyError:
            raise ValueError(f'unknown precision {p!r}; use one of {sorted(_PREC_NAMES)}')
    return int(p)


def prec_name(p: int) -> str:
    return {PREC_FP32: 'fp32', PREC_BF16X3: 'bf16x3', PREC_BF16: 'bf16'}[int(p)]


def set_default_precision(p):
    """'fp32' (CUDA-core exact order), 'bf16x3' (tcgen05 3-pass split, default) or 'bf16'."""
    global _default_prec
    _default_prec = parse_prec(p)


def get_default_precision() -> int:
    return _default_prec


def make_desc(nd: int, B: int, Cin: int, Cout: int, inp: Sequence[int], k: Sequence[int], stride: Sequence[int],
              pad_lo: Sequence[int], pad_hi: Sequence[int], transposed: bool) -> HebbDesc:
    def three(v, fill):
        v = list(v)
        return [fill] * (3 - len(v)) + v
    d = HebbDesc()
    d.nd, d.B, d.Cin, d.Cout = nd, B, Cin, Cout
    d.inp[:] = three(inp, 1)
    d.k[:] = three(k, 1)
    d.stride[:] = three(stride, 1)
    d.pad_lo[:] = three(pad_lo, 0)
    d.pad_hi[:] = three(pad_hi, 0)
    d.transposed = 1 if transposed else 0
    return d


def out_shape(desc: HebbDesc):
    out = (ctypes.c_int32 * 3)()
    check(load().hebb_out_shape(ctypes.byref(desc), ctypes.byref(out)), 'out_shape')
    return list(out)[3 - desc.nd:]


def workspace_bytes(desc: HebbDesc, prec: int) -> int:
    n = ctypes.c_size_t(0)
    check(load().hebb_workspace_bytes(ctypes.byref(desc), prec, ctypes.byref(n)), 'workspace_bytes')
    return int(n.value)


# One scratch buffer per device, shared by every layer (layers of a network run back to back on
# one stream, so their scratch never overlaps in time).  Grown on demand, never shrunk.
_ws = {}


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            # a stream may still be using the old buffer; let the caching allocator order it
            buf.record_stream(torch.cuda.current_stream(device))
        buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def release_workspaces():
    _ws.clear()


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor: the sm_100 Hebbian library has no CPU fallback')
    if t.dtype != torch.float32:
        raise RuntimeError(f'{name} must be float32 (got {t.dtype})')


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _device_guard(argidx):
    """Run the wrapped launch with the CUDA device of its tensor argument current (like torch's own ops): the C
    library launches on the current device with the stream handle of the tensor's device."""
    def deco(fn):
        import functools

        @functools.wraps(fn)
        def wrapped(*a, **k):
            t = a[argidx]
            if isinstance(t, (list, tuple)):
                t = t[0] if len(t) else None
            if t is None or not getattr(t, 'is_cuda', False) or t.device.index == torch.cuda.current_device():
                return fn(*a, **k)
            with torch.cuda.device(t.device):
                return fn(*a, **k)
        return wrapped
    return deco


@_device_guard(1)
def conv_step(desc: HebbDesc, x, W, bias, kinv: float, y, winner, delta_w, flags: int, prec: int):
    """Launch hebb_conv_swta_step / hebb_convT_swta_step (by desc.transposed) on the current stream."""
    lib = load()
    _require_cuda(x, 'x'); _require_cuda(W, 'weight')
    nbytes = workspace_bytes(desc, prec)
    ws = workspace(x.device, nbytes)
    fn = lib.hebb_convT_swta_step if desc.transposed else lib.hebb_conv_swta_step
    st = fn(ctypes.byref(desc), x.data_ptr(), W.data_ptr(), bias.data_ptr() if bias is not None else None,
            float(kinv), y.data_ptr(), winner.data_ptr() if winner is not None else None,
            delta_w.data_ptr() if delta_w is not None else None, ws.data_ptr(), ws.numel(),
            int(flags), int(prec), _stream_ptr(x.device))
    check(st, 'conv_swta_step')


@_device_guard(1)
def conv_step_stats(desc: HebbDesc, x, W, bias, kinv: float, y, winner, delta_w, flags: int, prec: int):
    """hebb_conv_swta_step_stats: the step plus the BatchNorm statistics of y.  Returns the [Cout, 2] float64
    tensor (sum, sum of squares per channel) or None when the library did not produce them."""
    lib = load()
    _require_cuda(x, 'x'); _require_cuda(W, 'weight')
    ws = workspace(x.device, workspace_bytes(desc, prec))
    stats = torch.empty((desc.Cout, 2), dtype=torch.float64, device=x.device)
    written = ctypes.c_int32(0)
    st = lib.hebb_conv_swta_step_stats(ctypes.byref(desc), x.data_ptr(), W.data_ptr(), bias.data_ptr() if bias is not None else None,
                                       float(kinv), y.data_ptr(), winner.data_ptr() if winner is not None else None,
                                       delta_w.data_ptr() if delta_w is not None else None, ws.data_ptr(), ws.numel(),
                                       int(flags), int(prec), stats.data_ptr(), ctypes.byref(written), _stream_ptr(x.device))
    check(st, 'conv_swta_step_stats')
    return stats if written.value else None


@_device_guard(0)
def bn_act_from_stats(y, stats, gamma, beta, running_mean, running_var, eps, momentum, slope, drop_p: float = 0.0):
    """BatchNorm(train) + activation of y from precomputed per-channel (sum, sum of squares) (hebb_bn_act_from_stats);
    drop_p > 0 folds the nn.Dropout(drop_p) that follows into the same pass (hebb_bn_act_from_stats_dropout)."""
    _require_cuda(y, 'input')
    B, C = y.shape[0], y.shape[1]
    S = y.numel() // (B * C)
    out = torch.empty_like(y)
    ws = workspace(y.device, C * 8 + 64)
    ptr = lambda t: t.data_ptr() if t is not None else None
    if drop_p > 0.0:
        st = dropout_state(y.device)
        check(load().hebb_bn_act_from_stats_dropout(y.data_ptr(), out.data_ptr(), stats.data_ptr(), ptr(gamma), ptr(beta),
                                                    ptr(running_mean), ptr(running_var), B, C, S, float(eps), float(momentum),
                                                    float(slope), float(drop_p), st.data_ptr(), ws.data_ptr(), ws.numel(),
                                                    _stream_ptr(y.device)), 'bn_act_from_stats_dropout')
        return out
    check(load().hebb_bn_act_from_stats(y.data_ptr(), out.data_ptr(), stats.data_ptr(), ptr(gamma), ptr(beta), ptr(running_mean),
                                        ptr(running_var), B, C, S, float(eps), float(momentum), float(slope), ws.data_ptr(),
                                        ws.numel(), _stream_ptr(y.device)), 'bn_act_from_stats')
    return out


@_device_guard(1)
def conv_wgrad(desc: HebbDesc, x, grad_y, prec: int, gy_channels: int = 0, channels_last: bool = False):
    """grad_w[Cout][Cin][taps] of a stride-1 convolution on the tcgen05 contraction kernel (hebb_conv_wgrad).
    x / grad_y: dense NCHW (default) or dense channels_last storage.  Returns None when the layer is outside
    the tensor-core planner (the caller then uses ATen)."""
    if desc.transposed or prec == PREC_FP32 or wgrad_path(desc, prec) <= 0:
        return None
    taps = desc.k[0] * desc.k[1] * desc.k[2]
    if channels_last and desc.Cin <= 4 and taps > 1:
        return None
    _require_cuda(x, 'x'); _require_cuda(grad_y, 'grad_y')
    gw = torch.zeros((desc.Cout, desc.Cin, taps), dtype=torch.float32, device=x.device)
    ws = workspace(x.device, workspace_bytes(desc, prec))
    check(load().hebb_conv_wgrad(ctypes.byref(desc), x.data_ptr(), grad_y.data_ptr(), gw.data_ptr(), int(gy_channels),
                                 1 if channels_last else 0, ws.data_ptr(), ws.numel(), int(prec),
                                 _stream_ptr(x.device)), 'conv_wgrad')
    return gw


@_device_guard(0)
def wnorm(W: torch.Tensor, rows: int, row_stride: int, mid: int, mid_stride: int, inner: int,
          want_inv: bool = False):
    """normalize() over the raw storage addressing described in hebb_sm100.h."""
    _require_cuda(W, 'weight')
    out = torch.empty_strided(W.shape, W.stride(), dtype=W.dtype, device=W.device)
    inv = torch.empty(rows, dtype=torch.float32, device=W.device) if want_inv else None
    check(load().hebb_wnorm(W.data_ptr(), out.data_ptr(), inv.data_ptr() if inv is not None else None,
                            rows, row_stride, mid, mid_stride, inner, _stream_ptr(W.device)), 'wnorm')
    return (out, inv) if want_inv else out


@_device_guard(1)
def local_update_multi(grads, dws, alphas, has_grad):
    """grad_i = (1-a_i) grad_i - a_i dw_i  (or -a_i dw_i), dw_i = 0, for all i in ONE launch."""
    n = len(dws)
    if n == 0:
        return
    lib = load()
    vp = ctypes.c_void_p
    g = (vp * n)(*[t.data_ptr() for t in grads])
    d = (vp * n)(*[t.data_ptr() for t in dws])
    ne = (ctypes.c_int64 * n)(*[t.numel() for t in dws])
    al = (ctypes.c_float * n)(*[float(a) for a in alphas])
    hg = (ctypes.c_int32 * n)(*[1 if h else 0 for h in has_grad])
    check(lib.hebb_local_update_multi(n, g, d, ne, al, hg, _stream_ptr(dws[0].device)), 'local_update_multi')


PLAN_FIELDS = ['MB', 'f_SEGLEN', 'XST', 'WST', 'NACC', 'f_tmem', 'f_tiles', 'f_smem', 'by_kh', 'CM', 'CN', 'BLK', 'ST',
               'd_SEGLEN', 'ngrp', 'n_cin', 'n_cout', 'PS', 'blocks', 'd_tmem', 'd_smem', 'd_HL', 'ws_MiB', 'stackM', 'stackN', 'CT', 'n_ct', 'nrep', 'WG', 'reuse', 'rhalo', 'rsw', 'rs_BLK', 'rs_ST', 'rs_smem', 'rs_tmem', 'rs_PS', 'rs_stackM']


def plan(desc: HebbDesc, prec: int):
    out = (ctypes.c_int * 48)()
    n = load().hebb_debug_plan(ctypes.byref(desc), int(prec), out, 48)
    return dict(zip(PLAN_FIELDS, list(out)[:n])) if n else None


@_device_guard(0)
def bn_act_train(y, gamma, beta, running_mean, running_var, eps, momentum, slope, out=None, drop_p: float = 0.0):
    """BatchNorm(train) + (Leaky)ReLU on a contiguous [B, C, *spatial] fp32 CUDA tensor (hebb_bn_act_train); drop_p > 0
    folds the nn.Dropout(drop_p) that follows into the same pass (hebb_bn_act_train_dropout)."""
    _require_cuda(y, 'input')
    B, C = y.shape[0], y.shape[1]
    S = y.numel() // (B * C)
    if out is None:
        out = torch.empty_like(y)
    ws = workspace(y.device, C * 24 + 64)
    ptr = lambda t: t.data_ptr() if t is not None else None
    if drop_p > 0.0:
        st = dropout_state(y.device)
        check(load().hebb_bn_act_train_dropout(y.data_ptr(), out.data_ptr(), ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var),
                                               B, C, S, float(eps), float(momentum), float(slope), float(drop_p), st.data_ptr(),
                                               ws.data_ptr(), ws.numel(), _stream_ptr(y.device)), 'bn_act_train_dropout')
        return out
    check(load().hebb_bn_act_train(y.data_ptr(), out.data_ptr(), ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var),
                                   B, C, S, float(eps), float(momentum), float(slope), ws.data_ptr(), ws.numel(),
                                   _stream_ptr(y.device)), 'bn_act_train')
    return out


@_device_guard(0)
def upsample2x_bilinear(x):
    _require_cuda(x, 'input')
    B, C, H, W = x.shape
    out = torch.empty((B, C, 2 * H, 2 * W), dtype=x.dtype, device=x.device)
    check(load().hebb_upsample2x_bilinear(x.data_ptr(), out.data_ptr(), B * C, H, W, _stream_ptr(x.device)), 'upsample2x')
    return out


@_device_guard(0)
def maxpool2x(x):
    """max_pool{2,3}d(kernel_size=2, stride=2) of a contiguous fp32 CUDA tensor (hebb_maxpool2x)."""
    _require_cuda(x, 'input')
    if x.dim() == 4:
        B, C, H, W = x.shape
        D, pd, out_shape = 1, 0, (B, C, H // 2, W // 2)
    else:
        B, C, D, H, W = x.shape
        pd, out_shape = 1, (B, C, D // 2, H // 2, W // 2)
    out = torch.empty(out_shape, dtype=x.dtype, device=x.device)
    check(load().hebb_maxpool2x(x.data_ptr(), out.data_ptr(), B * C, D, H, W, pd, _stream_ptr(x.device)), 'maxpool2x')
    return out


def _dense_channel_inner(t):
    """(inner) such that element i of t's dense storage belongs to channel (i / inner) % C, or None."""
    nd = t.dim() - 2
    if t.is_contiguous():
        return t[0, 0].numel()
    cl = torch.channels_last if nd == 2 else (torch.channels_last_3d if nd == 3 else None)
    if cl is not None and t.is_contiguous(memory_format=cl):
        return 1
    return None


_dropout_state = {}


def dropout_state(device, reseed: bool = False):
    """The device-resident Philox state {seed, launches so far} of hebb_bias_relu_dropout_state on `device`.  The seed
    is drawn once from torch's CPU generator (so torch.manual_seed() before the first use, or reseed=True after a
    later manual_seed(), governs the masks); the launch counter lives on the device, which is what lets a captured
    step draw a new mask on every replay."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    st = _dropout_state.get(key)
    if st is None or reseed:
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        new = torch.tensor([seed, 0], dtype=torch.int64, device=torch.device('cuda', key))
        if st is None:
            _dropout_state[key] = st = new
        else:
            st.copy_(new)                          # keep the address: a captured graph may hold it
    return st


@_device_guard(0)
def bias_relu_dropout(z, bias, p: float, seed=None):
    """(out, mask) = hebb_bias_relu_dropout on z's dense storage (NCHW or channels_last; out keeps z's layout).
    seed=None takes the stream from dropout_state(z.device) (no host value in the launch: graph-capturable)."""
    _require_cuda(z, 'input')
    inner = _dense_channel_inner(z)
    if inner is None:
        raise RuntimeError('bias_relu_dropout needs a dense NCHW or channels_last tensor')
    out = torch.empty_like(z)                      # preserves the memory format
    mask = torch.empty_like(z, dtype=torch.uint8)
    if seed is None:
        st = dropout_state(z.device)
        check(load().hebb_bias_relu_dropout_state(z.data_ptr(), bias.data_ptr(), out.data_ptr(), mask.data_ptr(), z.numel(),
                                                  z.shape[1], inner, float(p), st.data_ptr(), _stream_ptr(z.device)),
              'bias_relu_dropout_state')
    else:
        check(load().hebb_bias_relu_dropout(z.data_ptr(), bias.data_ptr(), out.data_ptr(), mask.data_ptr(), z.numel(), z.shape[1],
                                            inner, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, _stream_ptr(z.device)), 'bias_relu_dropout')
    return out, mask


@_device_guard(0)
def mask_scale(gout, mask, scale: float):
    gz = torch.empty_like(mask, dtype=torch.float32)
    check(load().hebb_mask_scale(gout.data_ptr(), mask.data_ptr(), gz.data_ptr(), gout.numel(), float(scale),
                                 _stream_ptr(gout.device)), 'mask_scale')
    return gz


_sm_count = {}


@_device_guard(0)
def mask_scale_gb(gout, mask, scale: float):
    """(gz, gb): hebb_mask_scale plus the per-channel sums of gz in the same pass, for dense channels_last tensors whose
    channel count is a power of two (4..1024); None when the tensors do not qualify (the caller then sums gz itself)."""
    C = mask.shape[1]
    if (_dense_channel_inner(mask) != 1 or gout.stride() != mask.stride() or C < 4 or C > 1024 or (C & (C - 1))
            or gout.dtype != torch.float32):
        return None
    dev = mask.device
    sms = _sm_count.get(dev.index)
    if sms is None:
        sms = _sm_count[dev.index] = torch.cuda.get_device_properties(dev).multi_processor_count
    rows = sms * 8
    gz = torch.empty_like(mask, dtype=torch.float32)
    gb = torch.empty(C, dtype=torch.float32, device=dev)
    partial = torch.empty(rows * C, dtype=torch.float32, device=dev)
    check(load().hebb_mask_scale_gb(gout.data_ptr(), mask.data_ptr(), gz.data_ptr(), gb.data_ptr(), gout.numel(), C, float(scale),
                                    partial.data_ptr(), rows, _stream_ptr(dev)), 'mask_scale_gb')
    return gz, gb


def launch_count() -> int:
    return int(load().hebb_debug_launch_count())


PATH_SIMT, PATH_TC, PATH_FUSED = 0, 1, 2
FUSED_PLAN_FIELDS = ['TH', 'TW', 'pitch', 'tiles', 'blocks_per_tile', 'x_rows', 'smem', 'tmem_cols', 'grid', 'stages', 'r_slots']


def layer_path(desc: HebbDesc, prec: int, flags: int = 0) -> int:
    """Which kernels a step of this layer runs on: PATH_SIMT (fp32 CUDA cores), PATH_TC (pack + forward + update
    tcgen05 kernels) or PATH_FUSED (the one-kernel small-channel path)."""
    return int(load().hebb_layer_path(ctypes.byref(desc), int(prec), int(flags)))


def wgrad_path(desc: HebbDesc, prec: int) -> int:
    """Which kernels conv_wgrad runs on: 0 none, PATH_TC, or PATH_FUSED (the fused kernel with dL/dy as the responses)."""
    return int(load().hebb_wgrad_path(ctypes.byref(desc), int(prec)))


def fused_plan(desc: HebbDesc):
    out = (ctypes.c_int * 16)()
    n = load().hebb_debug_fused_plan(ctypes.byref(desc), out, 16)
    return dict(zip(FUSED_PLAN_FIELDS, list(out)[:n])) if n else None


def uses_tensor_cores(desc: HebbDesc, prec: int) -> bool:
    return bool(load().hebb_uses_tensor_cores(ctypes.byref(desc), int(prec)))


def query():
    lib = load()
    a, b, c = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    st = lib.hebb_query(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return st, a.value, b.value, c.value
