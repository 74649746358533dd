"""Dimension-agnostic implementation behind HebbianConv{2,3}d / HebbianConvTranspose{2,3}d.

The public classes in hebb.py / hebb3d.py keep the reference's constructor signatures,
attributes, state_dict keys and method names (reference hebb/hebb.py:16-277,
hebb/hebb3d.py:15-305); everything numerical is a call into libhebb_sm100.so.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native


def normalize(x, dim=None):
    """x / ||x||_2 over `dim` with zero norms mapped to 1 (reference hebb/hebb.py:10-13).

    CUDA float32 tensors whose reduced dims are every dim but the first are normalised by the
    hebb_wnorm kernel; anything else (other dims, other dtypes, CPU tensors — this function is
    also a plain utility in the reference) is evaluated with ordinary tensor ops.
    """
    nd = x.dim()
    dims = tuple(range(nd)) if dim is None else tuple(d % nd for d in (dim if isinstance(dim, (tuple, list)) else (dim,)))
    if (x.is_cuda and x.dtype == torch.float32 and nd >= 2 and dims == tuple(range(1, nd))
            and not (torch.is_grad_enabled() and x.requires_grad)):
        lay = _dense_layout(x)
        if lay is not None:
            return _native.wnorm(x.detach(), *lay)
    nrm = (x ** 2).sum(dim=dims, keepdim=True) ** 0.5
    nrm = torch.where(nrm == 0, torch.ones_like(nrm), nrm)
    return x / nrm


def _dense_layout(w: torch.Tensor):
    """(rows,row_stride,mid,mid_stride,inner) for hebb_wnorm, or None if w is neither contiguous
    nor the transposed (dim0<->dim1) view of a contiguous buffer."""
    if w.is_contiguous():
        inner = w[0].numel()
        return (w.shape[0], inner, 1, 0, inner)
    base = w.transpose(0, 1)
    if base.is_contiguous() and w.storage_offset() == base.storage_offset():
        taps = math.prod(w.shape[2:]) if w.dim() > 2 else 1
        return (w.shape[0], taps, w.shape[1], w.shape[0] * taps, taps)
    return None


def _ntuple(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


class _HebbFn(torch.autograd.Function):
    """Forward = the sm_100 kernels.  Backward (only reached when back-prop gradients are wanted, i.e.
    alpha < 1 or a trainable layer upstream; reference hebb/hebb.py:185-191 via loss.backward()):
    stride-1 convolutions the tcgen05 planner takes run natively as well -- dL/dW on the plasticity
    contraction kernel with dL/dy in place of the responses (hebb_conv_wgrad), dL/dx as a forward
    convolution of dL/dy with the flipped, transposed filters -- everything else (transposed layers, strided
    or odd shapes, fp32 mode) differentiates the same formula with stock ATen ops.  SURVEY.md section 8(f) row 3."""

    @staticmethod
    def forward(ctx, x, weight, bias, layer, update):
        ctx.layer = layer
        ctx.save_for_backward(x, weight, bias)
        return layer._launch(x, weight, bias, update)

    @staticmethod
    def backward(ctx, gy):
        x, weight, bias = ctx.saved_tensors
        layer = ctx.layer
        need = ctx.needs_input_grad
        native = layer._native_backward(x, weight, bias, gy, need)
        if native is not None:
            return native[0], native[1], native[2], None, None
        with torch.enable_grad():
            xx = x.detach().requires_grad_(need[0])
            ww = weight.detach().requires_grad_(need[1])
            bb = bias.detach().requires_grad_(need[2]) if bias is not None else None
            yy = layer._aten_formula(xx, ww, bb)
            wanted = [t for t, n in ((xx, need[0]), (ww, need[1]), (bb, need[2])) if n and t is not None]
            got = list(torch.autograd.grad(yy, wanted, gy, allow_unused=True))
        out = []
        for t, n in ((xx, need[0]), (ww, need[1]), (bb, need[2])):
            out.append(got.pop(0) if (n and t is not None) else None)
        return out[0], out[1], out[2], None, None


class _HebbianConvNd(nn.Module):
    """Shared body.  Sub-classes set `_nd` and `_transposed`."""

    MODE_SWTA = 'swta'
    MODE_HPCA = 'hpca'
    MODE_CONTRASTIVE = 'contrastive'
    _nd = 2
    _transposed = False

    def _setup(self, in_channels, out_channels, kernel_size, stride, padding, bias, w_nrm, act,
               mode, k, patchwise, contrast, uniformity, alpha):
        nd = self._nd
        self.mode = mode
        self.out_channels = out_channels
        self.in_channels = in_channels
        self.kernel_size = _ntuple(kernel_size, nd)
        self.padding = padding
        self.stride = _ntuple(stride, nd)
        self.weight = nn.Parameter(torch.empty((out_channels, in_channels, *self.kernel_size)), requires_grad=True)
        nn.init.xavier_normal_(self.weight)
        self.w_nrm = w_nrm
        self.bias = nn.Parameter(torch.zeros(out_channels), requires_grad=bias)
        self.act = act
        self.register_buffer('delta_w', torch.zeros_like(self.weight))
        self.k = k
        self.patchwise = patchwise
        self.contrast = contrast
        self.uniformity = uniformity
        self.alpha = alpha
        if self._transposed:
            # the reference keeps (Cin, Cout, k...) VIEWS of (Cout, Cin, k...) storage (hebb.py:222-224)
            with torch.no_grad():
                self.weight.transpose_(0, 1)
                self.delta_w.transpose_(0, 1)
        # --- not part of the reference API ---
        self.prec = None              # None -> library default (hebb.set_precision / HEBB_PREC)
        self.record_winners = False   # when True the last forward's argmax map is kept in .winners
        self.winners = None
        self._desc_cache = {}
        self._emit_y_stats = False     # set by hebb.fused.fuse_norm_act when a fused BatchNorm consumes the output
        self._y_stats = None           # (y, [Cout, 2] float64 sums) of the last forward, if produced

    # ------------------------------------------------------------------ geometry
    def _pad_list(self):
        nd = self._nd
        p = self.padding
        if isinstance(p, int):
            return [p] * (2 * nd)
        p = list(p)
        if len(p) == nd:
            out = []
            for v in p:
                out += [v, v]
            return out
        return p

    def _pad_lo_hi(self):
        """(lo, hi) per spatial axis in (D,)H,W order, as F.pad would apply self._pad_list()."""
        nd = self._nd
        lst = self._pad_list()
        if len(lst) % 2 or len(lst) > 2 * nd:
            raise RuntimeError(f'padding {self.padding!r} pads non-spatial dims; unsupported by the sm_100 Hebbian layer')
        lo, hi = [0] * nd, [0] * nd
        for i in range(len(lst) // 2):          # pair i pads dim -(i+1)
            lo[nd - 1 - i], hi[nd - 1 - i] = int(lst[2 * i]), int(lst[2 * i + 1])
        if min(lo + hi) < 0:
            raise RuntimeError('negative padding is not supported by the sm_100 Hebbian layer')
        return lo, hi

    def _desc(self, x_shape, pad):
        key = (tuple(x_shape), pad)
        d = self._desc_cache.get(key)
        if d is None:
            lo, hi = self._pad_lo_hi() if pad else ([0] * self._nd, [0] * self._nd)
            d = _native.make_desc(self._nd, x_shape[0], self.in_channels, self.out_channels, x_shape[2:],
                                  self.kernel_size, self.stride, lo, hi, self._transposed)
            d._out = _native.out_shape(d)
            if len(self._desc_cache) > 16:
                self._desc_cache.clear()
            self._desc_cache[key] = d
        return d

    # ------------------------------------------------------------------ storage helpers
    def _raw(self, t):
        """The contiguous (Cout, Cin, k...) tensor that shares (or mirrors) t's storage."""
        if not self._transposed:
            return (t if t.is_contiguous() else None)
        b = t.transpose(0, 1)
        return b if b.is_contiguous() else None

    # ------------------------------------------------------------------ kernels
    def _launch(self, x, weight, bias, update, pad=True, w_nrm=None):
        if x.dim() != self._nd + 2:
            raise RuntimeError(f'expected a {self._nd + 2}-D input, got {tuple(x.shape)}')
        if not x.is_cuda:
            raise RuntimeError('Hebbian layers of this package run on CUDA (sm_100) only: move the model and '
                               'its inputs to the GPU; there is no CPU fallback')
        x = x.detach()
        if x.dtype != torch.float32:
            raise RuntimeError(f'input must be float32, got {x.dtype}')
        if not x.is_contiguous():
            x = x.contiguous()
        w = self._raw(weight.detach())
        tmp_w = w is None
        if tmp_w:   # unusual strides (e.g. after a user re-assignment): work on a dense copy
            w = (weight.detach().transpose(0, 1) if self._transposed else weight.detach()).contiguous()
        desc = self._desc(x.shape, pad)
        y = torch.empty((x.shape[0], self.out_channels, *desc._out), dtype=torch.float32, device=x.device)
        win = None
        if self.record_winners:
            win = torch.empty((x.shape[0], *desc._out), dtype=torch.int32, device=x.device)
        flags = _native.F_WNRM if (self.w_nrm if w_nrm is None else w_nrm) else 0
        dw = None
        tmp_dw = None
        if update:
            flags |= _native.F_UPDATE
            if self.mode == self.MODE_HPCA:
                flags |= _native.F_RULE_HPCA
            dw = self._raw(self.delta_w)
            if dw is None:
                tmp_dw = torch.zeros_like(w)
                dw = tmp_dw
        b = bias.detach() if bias is not None else None
        self._y_stats = None
        if self._emit_y_stats and not self._transposed and pad:
            # a fused BatchNorm follows (hebb.fused.fuse_norm_act): its statistics ride along in the forward epilogue
            stats = _native.conv_step_stats(desc, x, w, b, float(self.k), y, win, dw, flags, _native.parse_prec(self.prec))
            if stats is not None:
                self._y_stats = (y, stats)
        else:
            _native.conv_step(desc, x, w, b, float(self.k), y, win, dw, flags, _native.parse_prec(self.prec))
        if tmp_dw is not None:
            self.delta_w += tmp_dw.transpose(0, 1) if self._transposed else tmp_dw
        self.winners = win
        return y

    def _native_backward(self, x, weight, bias, gy, need):
        """(dL/dx, dL/dW, dL/db) on the sm_100 kernels, or None when this layer / shape is not covered."""
        import os
        if os.environ.get('HEBB_ATEN_BACKWARD') == '1' or self._transposed or not x.is_cuda:
            return None
        nd = self._nd
        prec = _native.parse_prec(self.prec)
        if prec == _native.PREC_FP32 or any(s != 1 for s in _ntuple(self.stride, nd)):
            return None
        w_raw = self._raw(weight.detach())
        if w_raw is None or x.dtype != torch.float32:
            return None
        x = x.detach().contiguous()
        gy = gy.detach().contiguous()
        desc = self._desc(x.shape, True)
        lo, hi = self._pad_lo_hi()
        ks = _ntuple(self.kernel_size, nd)
        gx = gw = gb = None
        if need[0]:
            lo2 = [k - 1 - p for k, p in zip(ks, lo)]
            hi2 = [k - 1 - p for k, p in zip(ks, hi)]
            if min(lo2 + hi2) < 0:
                return None
            d2 = _native.make_desc(nd, x.shape[0], self.out_channels, self.in_channels, gy.shape[2:], ks,
                                   (1,) * nd, lo2, hi2, False)
            if not _native.uses_tensor_cores(d2, prec):
                return None
        if need[1] and not _native.uses_tensor_cores(desc, prec):
            return None
        if need[0]:
            wn = normalize(weight.detach(), dim=tuple(range(1, nd + 2))) if self.w_nrm else weight.detach()
            wf = wn.flip(tuple(range(2, nd + 2))).transpose(0, 1).contiguous()
            gx = torch.empty_like(x)
            _native.conv_step(d2, gy, wf, None, 1.0, gx, None, None, 0, prec)
        if need[1]:
            gwn = _native.conv_wgrad(desc, x, gy, prec).reshape(weight.shape)
            if self.w_nrm:          # chain rule through W / |W| on the (small) weight tensor
                with torch.enable_grad():
                    ww = weight.detach().requires_grad_(True)
                    nrm = (ww ** 2).sum(dim=tuple(range(1, ww.dim())), keepdim=True) ** 0.5
                    wn2 = ww / torch.where(nrm == 0, torch.ones_like(nrm), nrm)
                    gw, = torch.autograd.grad(wn2, ww, gwn)
            else:
                gw = gwn
        if need[2] and bias is not None:
            gb = gy.sum(dim=(0, *range(2, nd + 2)))
        return gx, gw, gb

    def _aten_formula(self, x, w, b):
        """The reference formula with stock ops; used ONLY to differentiate (see _HebbFn.backward)."""
        x = F.pad(x, self._pad_list())
        if self.w_nrm:
            nrm = (w ** 2).sum(dim=tuple(range(1, w.dim())), keepdim=True) ** 0.5
            w = w / torch.where(nrm == 0, torch.ones_like(nrm), nrm)
        if self._transposed:
            op = torch.conv_transpose2d if self._nd == 2 else torch.conv_transpose3d
        else:
            op = torch.conv2d if self._nd == 2 else torch.conv3d
        return op(x, w, bias=b, stride=self.stride)

    # ------------------------------------------------------------------ reference API
    def apply_weights(self, x, w):
        """conv / conv_transpose of an already padded x with the given weights (no normalisation)."""
        return self._launch(x, w, self.bias, update=False, pad=False, w_nrm=False)

    def compute_activation(self, x):
        """y = act(conv(x, W/|W|)) for an already padded x."""
        return self.act(self._launch(x, self.weight, self.bias, update=False, pad=False))

    def pad(self, x):
        return F.pad(x, self._pad_list())

    def _check_mode(self):
        valid = [self.MODE_SWTA, self.MODE_HPCA, self.MODE_CONTRASTIVE]
        if self._transposed:
            valid += [self.MODE_SWTA_T, self.MODE_HPCA_T]
        if self.mode not in valid:
            raise NotImplementedError("Learning mode {} unavailable for {} layer".format(self.mode, self.__class__.__name__))
        native = self.MODE_SWTA_T if self._transposed else self.MODE_SWTA
        if self._transposed and self.mode == self.MODE_HPCA_T:
            return                      # hebb.py:266-277: see _hpca_t_update
        if self.mode == self.MODE_HPCA and self.patchwise:
            return                      # HPCA: tcgen05 kernels for plain convs; transposed layers (the conv rule
                                        # with x and y exchanged, hebb.py:243-246) on the fp32 CUDA-core kernels
        if self.mode == self.MODE_CONTRASTIVE:
            # Built for what the reference itself can execute: 2-D plain conv layers without uniformity weighting
            # (its 3-D twin calls unfold3d() with a zero stride, hebb3d.py:170, and the uniformity branch adds the
            # [Cout] bias to a 1-channel map, hebb.py:75,160 -- both raise RuntimeError in the reference).
            if self._nd != 2 or self._transposed or self.uniformity:
                raise NotImplementedError(
                    "Learning mode contrastive is available for HebbianConv2d with uniformity=False (the only "
                    "configuration the reference implementation executes without raising)")
            return
        if self.mode != native or not self.patchwise:
            raise NotImplementedError(
                "Learning mode {} (patchwise={}) of {} is not built into libhebb_sm100 yet; the sm_100 library "
                "implements the pretraining path ('{}', patchwise=True) and has no PyTorch fallback".format(
                    self.mode, self.patchwise, self.__class__.__name__, native))

    def _hpca_t_update(self, x, y):
        """mode 'hpca_t' of the transposed layers (hebb.py:266-277, hebb3d.py:291-305): for every kernel offset t the
        responses are the outputs y at offset t of each (non-overlapping) patch,
            delta_w[ci, co, t] += sum_p y_t[co, p] x[ci, p] - sum_t' sum_{co' <= co} (y_t' y_t'^T)[co, co'] W[ci, co', t']
        (the decay is summed over the offsets for patchwise=True, per offset otherwise).  A rule off the pretraining path
        (SURVEY 8f row 1): the forward runs on the sm_100 kernels, the update is a few batched GEMMs through torch on
        the same device.  The 3-D reference evaluates the triangular decay in chunks of 32 output channels
        (PARALLEL_CHANNELS, hebb3d.py:12,293-305); that is kept."""
        nd = self._nd
        ks, st = self.kernel_size, self.stride
        if tuple(ks) != tuple(st) or any(v != 0 for v in self._pad_list()):
            raise NotImplementedError('hpca_t is built for kernel_size == stride and padding 0 (the layers makehebbian produces)')
        with torch.no_grad():
            B, C = y.shape[0], y.shape[1]
            sp = x.shape[2:]
            taps = math.prod(ks)
            shape = [B, C]
            for n, k in zip(sp, ks):
                shape += [n, k]
            yv = y.detach().reshape(shape)
            k_axes = [3 + 2 * i for i in range(nd)]
            s_axes = [2 + 2 * i for i in range(nd)]
            r = yv.permute(*k_axes, 1, 0, *s_axes).reshape(taps, C, -1)                  # [taps, Cout, B * P]
            xf = x.detach().permute(0, *range(2, nd + 2), 1).reshape(-1, x.shape[1])     # [B * P, Cin]
            w = self.weight.detach()                                                      # (Cin, Cout, k...) view
            wp = w.permute(*range(2, nd + 2), 1, 0).reshape(taps, C, -1)                  # [taps, Cout, Cin]
            step = 32 if nd == 3 else C
            for c0 in range(0, C, step):
                c1 = min(C, c0 + step)
                ri = r[:, c0:c1]
                tri = torch.tril(torch.ones(c1 - c0, c1 - c0, device=x.device, dtype=x.dtype))
                dec = (ri.matmul(ri.transpose(-2, -1)) * tri.unsqueeze(0)).matmul(wp[:, c0:c1])
                if self.patchwise:
                    dec = dec.sum(dim=0, keepdim=True)
                upd = (ri.matmul(xf.unsqueeze(0)) - dec).permute(2, 1, 0)                 # [Cin, chunk, taps]
                self.delta_w[:, c0:c1] += upd.reshape(self.delta_w[:, c0:c1].shape)

    def _act_is_identity(self):
        return self.act is None or isinstance(self.act, nn.Identity)

    def _update_through_act(self, x, ya, pad=True):
        """Soft-WTA update for a layer with a non-Identity `act`: the reference feeds act(conv(x)) to the rule
        (hebb.py:80,87-90,107), so the responses are r = softmax_c(k * act(y)) and the fused epilogue (which sees the
        pre-activation y) cannot be used.  r is formed with element-wise torch ops; both contractions stay on the
        sm_100 kernels: delta_w += r X - (sum_p r) W  with r X = hebb_conv_wgrad(x, r)."""
        if self._transposed or self.mode != self.MODE_SWTA or not self.patchwise:
            raise NotImplementedError(
                'a non-Identity act is supported for plain convolutions in swta/patchwise mode only '
                '(the reference applies the plasticity rule to act(y): hebb.py:80,107)')
        prec = _native.parse_prec(self.prec)
        if prec == _native.PREC_FP32:
            prec = _native.PREC_BF16X3
        desc = self._desc(x.shape, pad)
        with torch.no_grad():
            r = (ya.detach() * float(self.k)).softmax(dim=1).contiguous()
            h = _native.conv_wgrad(desc, x.detach().contiguous(), r, prec)
            if h is None:
                raise NotImplementedError('this layer shape is outside the tensor-core planner; a non-Identity act is '
                                          'only supported on layers the tcgen05 kernels take')
            rs = r.sum(dim=tuple(i for i in range(r.dim()) if i != 1))
            w = self.weight.detach()
            self.delta_w += h.reshape(w.shape) - rs.view(-1, *([1] * (w.dim() - 1))) * w

    def forward(self, x):
        update = bool(self.training and self.alpha != 0)
        if update:
            self._check_mode()
        if update and self._transposed and self.mode == self.MODE_HPCA_T:
            y = self.act(self._forward_no_update(x))
            self._hpca_t_update(x, y)
            return y
        if update and not self._act_is_identity() and self.mode != self.MODE_CONTRASTIVE:
            y = self.act(self._forward_no_update(x))
            self._update_through_act(x, y)
            return y
        w = self.weight
        if self.alpha == 1:
            w = w.detach()       # (1 - alpha) * grad == 0 in local_update(): no need to back-prop into W
        b = self.bias
        contrastive = update and self.mode == self.MODE_CONTRASTIVE
        if contrastive:
            update = False               # the rule is a loss on the output, not a fused plasticity kernel
        if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or b.requires_grad):
            y = _HebbFn.apply(x, w, b, self, update)
        else:
            y = self._launch(x, w, b, update)
        if contrastive:
            self._contrastive_update(x)
        return self.act(y)

    def _forward_no_update(self, x):
        w = self.weight.detach() if self.alpha == 1 else self.weight
        b = self.bias
        if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or b.requires_grad):
            return _HebbFn.apply(x, w, b, self, False)
        return self._launch(x, w, b, False)

    def _contrastive_update(self, x):
        """hebb.py:143-172: delta_w += dL/dW of  L = sum [ -(S*y) + contrast * (S[perm]*y) ],  y = the layer
        output normalised over channels, S = its 3x3 box sum, perm = torch.randperm(batch).  The forward and the
        weight gradient run on the sm_100 kernels (through _HebbFn); the pixel-wise loss is a handful of
        element-wise torch ops.  As in the reference, L.backward() also deposits dL/dbias in bias.grad."""
        with torch.enable_grad():
            w, b = self.weight, self.bias
            leaves = [t for t in (w, b) if t.requires_grad]
            if w not in leaves:
                raise RuntimeError('contrastive learning needs weight.requires_grad=True (it is a gradient of a loss)')
            y = self.act(_HebbFn.apply(x.detach(), w, b, self, False))
            nrm = (y ** 2).sum(dim=1, keepdim=True) ** 0.5
            y = y / torch.where(nrm == 0, torch.ones_like(nrm), nrm)
            S = F.avg_pool2d(y, 3, stride=1, padding=1, count_include_pad=True) * 9.0
            idx = torch.randperm(y.size(0), device=y.device)
            L = (-(S * y) + self.contrast * S[idx] * y).sum()
            grads = torch.autograd.grad(L, leaves)
        with torch.no_grad():
            self.delta_w += grads[0]
            if len(leaves) > 1:
                b.grad = grads[1] if b.grad is None else b.grad + grads[1]

    def compute_update(self, x, y):
        """Accumulate the plasticity update for an already padded x into delta_w (y is recomputed
        on chip; the argument is accepted for signature compatibility)."""
        self._check_mode()
        if self._transposed and self.mode == self.MODE_HPCA_T:
            return self._hpca_t_update(x, y)
        if not self._act_is_identity() and self.mode != self.MODE_CONTRASTIVE:
            return self._update_through_act(x, y, pad=False)
        self._launch(x, self.weight, self.bias, update=True, pad=False)

    @torch.no_grad()
    def local_update(self):
        """weight.grad = (1 - alpha) * weight.grad - alpha * delta_w ; delta_w = 0  (hebb.py:174-192)."""
        dw = self.delta_w
        had = self.weight.grad is not None
        if not dw.is_cuda:
            raise RuntimeError('local_update(): the layer must live on a CUDA device (no CPU fallback)')
        if not had:
            self.weight.grad = torch.empty_like(dw)           # same (possibly transposed) strides as delta_w
        g = self.weight.grad
        if g.stride() != dw.stride() or _dense_layout(dw) is None:
            g = g.contiguous() if had else g
            new = torch.empty_like(dw, memory_format=torch.contiguous_format)
            d2 = dw.contiguous()
            if had:
                new.copy_(g)
            _native.local_update_multi([new], [d2], [self.alpha], [had])
            self.weight.grad = new
            dw.zero_()
            return
        _native.local_update_multi([g], [dw], [self.alpha], [had])
