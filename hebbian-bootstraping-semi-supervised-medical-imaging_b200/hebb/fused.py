"""Opt-in fusion of the bandwidth-bound ops around the Hebbian convolutions.

`fuse_norm_act(model)` rewrites, in place and without touching parameter names or state_dict keys,
every `HebbianConv -> BatchNorm{2,3}d -> ReLU/LeakyReLU` run inside an `nn.Sequential` so that the
BatchNorm(train) + activation pair runs as `hebb_bn_act_train` (one statistics pass, one
normalise+activate pass) — or, with `fuse_stats=True`, as `hebb_bn_act_from_stats` on the per-channel sums
the producing layer's forward epilogue hands back (`hebb_conv_swta_step_stats`: no statistics pass at all) —
and every `nn.Upsample(scale_factor=2, bilinear, align_corners=True)` so it runs as `hebb_upsample2x_bilinear`, and every 2x `nn.MaxPool{2,3}d` (kernel = stride = 2, no padding) so it runs
as `hebb_maxpool2x`; the stock convolutions a Hebbian network keeps for back-prop (makehebbian's `exclude` list)
have their `Conv -> ReLU -> Dropout` runs turned into convolution-without-bias + one `hebb_bias_relu_dropout`
pass (`fuse_head_act=True`; statistically the same dropout, its own Philox stream), and can get their weight
gradient from `hebb_conv_wgrad` (`head_wgrad=N`: convolutions with at most N filters, default 64 — the 2-class
output layer, whose cuDNN weight gradient costs 1.7 ms against 0.9 ms here, fp32-equivalent instead of TF32; the
16 -> 64 and 64 -> 32 layers of the 2-D head where the fused kernel takes them in weight-gradient mode, i.e. reads
x and dL/dy once per channel pass instead of packing them first; other wide layers stay with cuDNN).
Numerics follow torch (biased variance for normalisation, unbiased for the running estimate, momentum update, num_batches_tracked).  Anything the kernels do not cover —
eval mode, inputs or affine parameters that require grad, CPU tensors, cumulative-average momentum —
takes the stock torch path of the parent class, so the pass is always safe to apply.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native


def _fast_ok(x, mod):
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in mod.parameters(recurse=False))):
        return False
    return True


class _FusedBNActMixin:
    """Mixed into a BatchNorm instance by fuse_norm_act(); `_act_slope` is the fused activation."""

    _drop_p = 0.0          # > 0: the nn.Dropout that followed the activation lives in this module too (fuse_norm_act)

    def forward(self, x):
        if not (self.training and self.track_running_stats and self.momentum is not None and _fast_ok(x, self)):
            y = super().forward(x)
            s = self._act_slope
            y = y if s == 1.0 else (F.relu(y) if s == 0.0 else F.leaky_relu(y, s))
            return F.dropout(y, self._drop_p, self.training) if self._drop_p > 0.0 else y
        self._check_input_dim(x)
        if self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
        src = self.__dict__.get('_src_conv')
        src = src[0] if src else None
        held = getattr(src, '_y_stats', None) if src is not None else None
        if held is not None and held[0] is x:
            # the producing Hebbian layer already summed y and y^2 in its epilogue: skip the statistics pass
            src._y_stats = None
            return _native.bn_act_from_stats(x, held[1], self.weight, self.bias, self.running_mean, self.running_var,
                                             self.eps, self.momentum, self._act_slope, drop_p=self._drop_p)
        return _native.bn_act_train(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps,
                                    self.momentum, self._act_slope, drop_p=self._drop_p)


class FusedBatchNormAct2d(_FusedBNActMixin, nn.BatchNorm2d):
    pass


class FusedBatchNormAct3d(_FusedBNActMixin, nn.BatchNorm3d):
    pass


class FastUpsample2x(nn.Upsample):
    def forward(self, x):
        if x.dim() == 4 and _fast_ok(x, self):
            return _native.upsample2x_bilinear(x)
        return super().forward(x)


def _is_two(v, n):
    v = tuple(v) if isinstance(v, (tuple, list)) else (v,) * n
    return len(v) == n and all(int(i) == 2 for i in v)


def _is_zero(v, n):
    v = tuple(v) if isinstance(v, (tuple, list)) else (v,) * n
    return all(int(i) == 0 for i in v)


def _is_one(v, n):
    v = tuple(v) if isinstance(v, (tuple, list)) else (v,) * n
    return all(int(i) == 1 for i in v)


class _FastMaxPoolMixin:
    def forward(self, x):
        nd = x.dim() - 2
        if _fast_ok(x, self) and min(x.shape[2:]) >= 2:
            return _native.maxpool2x(x)
        return super().forward(x)


class FastMaxPool2d(_FastMaxPoolMixin, nn.MaxPool2d):
    pass


class FastMaxPool3d(_FastMaxPoolMixin, nn.MaxPool3d):
    pass


def _pool_is_2x(m, n):
    stride = m.stride if m.stride is not None else m.kernel_size
    return (_is_two(m.kernel_size, n) and _is_two(stride, n) and _is_zero(m.padding, n) and _is_one(m.dilation, n)
            and not m.ceil_mode and not m.return_indices)


class _ConvWgradFn(torch.autograd.Function):
    """A stock convolution whose WEIGHT gradient runs on the Hebbian contraction kernel (hebb_conv_wgrad): the
    back-prop layers a Hebbian network keeps (the `exclude` list of makehebbian: the segmentation head) have the
    tall-skinny weight-gradient shape -- a few dozen filters reduced over millions of pixels -- that the update
    kernel is built for and cuDNN serves poorly.  Forward and dL/dx stay with cuDNN."""

    @staticmethod
    def forward(ctx, x, weight, bias, mod):
        ctx.mod = mod
        ctx.save_for_backward(x, weight)
        return mod._conv_forward(x, weight, bias)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        mod = ctx.mod
        need = ctx.needs_input_grad
        nd = x.dim() - 2
        gx = gw = gb = None
        if need[1]:
            gw = mod._native_wgrad(x, gy)
        mask = (need[0], need[1] and gw is None, False)
        if mask[0] or mask[1]:
            r = torch.ops.aten.convolution_backward(gy, x, weight, None, list(mod.stride), list(mod.padding), list(mod.dilation),
                                                    False, [0] * nd, mod.groups, list(mask))
            gx = r[0] if mask[0] else None
            gw = r[1] if mask[1] else gw
        if need[2]:
            gb = gy.sum(dim=(0, *range(2, nd + 2)))
        return gx, gw, gb, None


class _FastWgradMixin:
    def _native_wgrad(self, x, gy):
        """dL/dW through hebb_conv_wgrad, or None (shape / layout / precision not covered -> ATen)."""
        nd = x.dim() - 2
        if not (x.is_cuda and x.dtype == torch.float32 and gy.dtype == torch.float32):
            return None
        if _native.get_default_precision() == _native.PREC_FP32:
            return None
        prec = _native.PREC_BF16X3             # fp32-equivalent split (cuDNN's own path here is TF32)
        cl_fmt = torch.channels_last if nd == 2 else torch.channels_last_3d
        # x and dL/dy must share one dense layout (NCHW or channels_last).  When they differ, the SMALLER tensor is
        # copied: the 2-class output layer gets its dL/dy in NCHW from the softmax backward while its saved
        # 32-channel input is channels_last -- converting x there cost 0.58 ms per step, converting dL/dy 0.03 ms
        def layout(t):
            if t.is_contiguous():
                return 'nchw'
            return 'cl' if t.is_contiguous(memory_format=cl_fmt) else None
        lx, lg = layout(x), layout(gy)
        if lg is None and lx is None:
            return None
        if lx == lg:
            want = lx
        elif lx is None or lg is None:
            want = lx or lg
        else:
            want = lx if x.numel() >= gy.numel() else lg
        cl = want == 'cl'
        if cl:
            x, gy = x.contiguous(memory_format=cl_fmt), gy.contiguous(memory_format=cl_fmt)
        else:
            x, gy = x.contiguous(), gy.contiguous()
        cout = self.out_channels
        cpad = (cout + 15) // 16 * 16          # the kernel works on multiples of 16 filters; the extra rows are zero
        desc = _native.make_desc(nd, x.shape[0], self.in_channels, cpad, x.shape[2:], self.kernel_size, (1,) * nd,
                                 self.padding, self.padding, False)
        if cout > 16 and _native.wgrad_path(desc, prec) != _native.PATH_FUSED:
            return None                        # wider layers only where the fused kernel takes them (else cuDNN is as fast)
        gw = _native.conv_wgrad(desc, x, gy, prec, gy_channels=cout, channels_last=cl)
        if gw is None:
            return None
        gw = gw[:cout].reshape(self.weight.shape)
        if self.weight.is_contiguous(memory_format=cl_fmt) and not self.weight.is_contiguous():
            gw = gw.contiguous(memory_format=cl_fmt)
        return gw

    def _fast_ok(self, x):
        return (self.padding_mode == 'zeros' and isinstance(self.padding, tuple) and self.groups == 1
                and all(s == 1 for s in self.stride) and all(d == 1 for d in self.dilation)
                and x.is_cuda and torch.is_grad_enabled() and self.weight.requires_grad)

    def forward(self, x):
        if not self._fast_ok(x):
            return super().forward(x)
        return _ConvWgradFn.apply(x, self.weight, self.bias, self)


class FastWgradConv2d(_FastWgradMixin, nn.Conv2d):
    pass


class FastWgradConv3d(_FastWgradMixin, nn.Conv3d):
    pass


class _BiasReluDropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, bias, p):
        out, mask = _native.bias_relu_dropout(z.detach(), bias.detach(), p)   # stream state on the device (_native.dropout_state)
        ctx.save_for_backward(mask)
        ctx.scale = 1.0 / (1.0 - p)
        return out

    @staticmethod
    def backward(ctx, gout):
        mask, = ctx.saved_tensors
        if gout.stride() != mask.stride():          # bring dL/dout into the layout the mask was written in
            cl = mask.dim() == 4 and _native._dense_channel_inner(mask) == 1
            gout = gout.contiguous(memory_format=torch.channels_last) if cl else gout.contiguous()
        if ctx.needs_input_grad[1]:
            both = _native.mask_scale_gb(gout, mask, ctx.scale)      # channels_last: bias gradient out of the same pass
            if both is not None:
                return both[0], both[1], None
        gz = _native.mask_scale(gout, mask, ctx.scale)
        gb = gz.sum(dim=(0, *range(2, gz.dim()))) if ctx.needs_input_grad[1] else None
        return gz, gb, None


class _NoBiasConvMixin:
    """The bias add moves into the fused activation kernel that follows (see FusedBiasReluDropout).  The convolution
    alone decides, per call, whether it leaves the bias out, and records that on the follower; the follower acts on
    the record (it never re-derives the decision from its own input, which may differ in dtype under autocast)."""

    def _omit_bias(self, x):
        follower = self.__dict__.get('_bias_follower')
        if follower is None:
            return False
        omit = follower[0]._wants_bias(x)
        follower[0]._bias_pending = omit
        return omit

    def forward(self, x):
        if self._omit_bias(x):
            return self._conv_forward(x, self.weight, None)
        return super().forward(x)


class NoBiasConv2d(_NoBiasConvMixin, nn.Conv2d):
    pass


class NoBiasFastWgradConv2d(_NoBiasConvMixin, _FastWgradMixin, nn.Conv2d):
    """Both: the bias add lives in the fused activation that follows, the weight gradient in hebb_conv_wgrad."""

    def forward(self, x):
        bias = None if self._omit_bias(x) else self.bias
        if self._fast_ok(x):
            return _ConvWgradFn.apply(x, self.weight, bias, self)
        return self._conv_forward(x, self.weight, bias)


class FusedBiasReluDropout(nn.ReLU):
    """Stands where the ReLU of a `Conv -> ReLU -> Dropout` run stood (the Dropout becomes an Identity): adds the
    convolution's bias, applies ReLU and dropout in one pass (hebb_bias_relu_dropout).  Parameter-free, so the
    state_dict is unchanged; eval mode, CPU tensors and odd layouts take the stock ops."""

    _bias_pending = False

    def _wants_bias(self, x):
        return self.training and x.is_cuda and x.dtype == torch.float32 and 0.0 <= self._p < 1.0

    def forward(self, z):
        conv = self.__dict__['_conv'][0]
        owes_bias, self._bias_pending = self._bias_pending, False
        if owes_bias and z.is_cuda and z.dtype == torch.float32 and _native._dense_channel_inner(z) is not None:
            return _BiasReluDropoutFn.apply(z, conv.bias, self._p)
        # stock path; the bias is added here exactly when the convolution reported that it left it out
        if owes_bias:
            z = z + conv.bias.to(z.dtype).view(1, -1, *([1] * (z.dim() - 2)))
        return F.dropout(F.relu(z), self._p, self.training)


def _slope_of(m):
    if type(m) is nn.ReLU:
        return 0.0
    if type(m) is nn.LeakyReLU:
        return float(m.negative_slope)
    return None


def fuse_norm_act(model: nn.Module, head_wgrad: int = 64, fuse_stats: bool = True, fuse_head_act: bool = True,
                  fuse_dropout: bool = True) -> nn.Module:
    n_bn = n_up = n_pool = n_head = n_act = n_drop = 0
    for mod in model.modules():
        if isinstance(mod, nn.Sequential):
            items = list(mod._modules.items())
            for i, (name, m) in enumerate(items):
                if type(m) in (nn.BatchNorm2d, nn.BatchNorm3d) and i > 0 and hasattr(items[i - 1][1], 'local_update'):
                    slope = _slope_of(items[i + 1][1]) if i + 1 < len(items) else None
                    m.__class__ = FusedBatchNormAct2d if type(m) is nn.BatchNorm2d else FusedBatchNormAct3d
                    m._act_slope = 1.0 if slope is None else slope
                    conv = items[i - 1][1]
                    if fuse_stats and type(getattr(conv, 'act', None)) is nn.Identity:
                        conv._emit_y_stats = True                    # statistics come out of the conv's epilogue
                        m.__dict__['_src_conv'] = (conv,)            # plain attribute, not a sub-module: no state_dict entry
                    if slope is not None:
                        mod._modules[items[i + 1][0]] = nn.Identity()     # activation now lives in the fused module
                        # ... and so does the Dropout behind it (Conv -> BN -> LeakyReLU -> Dropout -> Conv blocks)
                        if fuse_dropout and i + 2 < len(items) and type(items[i + 2][1]) is nn.Dropout:
                            d = items[i + 2][1]
                            if 0.0 < d.p < 1.0 and not d.inplace:
                                m._drop_p = float(d.p)
                                mod._modules[items[i + 2][0]] = nn.Identity()
                                n_drop += 1
                    n_bn += 1
        if fuse_head_act and isinstance(mod, nn.Sequential):
            items = list(mod._modules.items())
            for i in range(len(items) - 2):
                c, r, d = items[i][1], items[i + 1][1], items[i + 2][1]
                if (type(c) is nn.Conv2d and c.bias is not None and type(r) is nn.ReLU and type(d) is nn.Dropout
                        and 0.0 <= d.p < 1.0):
                    c.__class__ = NoBiasConv2d
                    act = FusedBiasReluDropout()
                    act._p = float(d.p)
                    act.__dict__['_conv'] = (c,)
                    c.__dict__['_bias_follower'] = (act,)
                    mod._modules[items[i + 1][0]] = act
                    mod._modules[items[i + 2][0]] = nn.Identity()
                    n_act += 1
        for name, m in list(mod._modules.items()):
            if type(m) is nn.Upsample and m.mode == 'bilinear' and m.align_corners and m.scale_factor in (2, 2.0, (2, 2), (2.0, 2.0)):
                m.__class__ = FastUpsample2x
                n_up += 1
            elif type(m) is nn.MaxPool2d and _pool_is_2x(m, 2):
                m.__class__ = FastMaxPool2d
                n_pool += 1
            elif type(m) is nn.MaxPool3d and _pool_is_2x(m, 3):
                m.__class__ = FastMaxPool3d
                n_pool += 1
            elif (head_wgrad and type(m) in (nn.Conv2d, nn.Conv3d, NoBiasConv2d) and m.weight.requires_grad
                  and m.out_channels <= int(head_wgrad)):
                m.__class__ = {nn.Conv2d: FastWgradConv2d, nn.Conv3d: FastWgradConv3d, NoBiasConv2d: NoBiasFastWgradConv2d}[type(m)]
                n_head += 1
    model._hebb_fused = dict(bn_act=n_bn, upsample=n_up, maxpool=n_pool, head_wgrad=n_head, bias_relu_dropout=n_act, bn_act_dropout=n_drop)
    return model
