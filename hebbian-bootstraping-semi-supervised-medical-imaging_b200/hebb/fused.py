"""Opt-in fusion of the bandwidth-bound ops around the Hebbian convolutions.

`fuse_norm_act(model)` rewrites, in place and without touching parameter names or state_dict keys,
every `HebbianConv -> BatchNorm{2,3}d -> ReLU/LeakyReLU` run inside an `nn.Sequential` so that the
BatchNorm(train) + activation pair runs as `hebb_bn_act_train` (one statistics pass, one
normalise+activate pass), and every `nn.Upsample(scale_factor=2, bilinear, align_corners=True)` so it
runs as `hebb_upsample2x_bilinear`, and every 2x `nn.MaxPool{2,3}d` (kernel = stride = 2, no padding) so it runs
as `hebb_maxpool2x`.  Numerics follow torch (biased variance for normalisation, unbiased
for the running estimate, momentum update, num_batches_tracked).  Anything the kernels do not cover —
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

    def forward(self, x):
        if not (self.training and self.track_running_stats and self.momentum is not None and _fast_ok(x, self)):
            y = super().forward(x)
            s = self._act_slope
            return y if s == 1.0 else (F.relu(y) if s == 0.0 else F.leaky_relu(y, s))
        self._check_input_dim(x)
        if self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
        return _native.bn_act_train(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps,
                                    self.momentum, self._act_slope)


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


def _slope_of(m):
    if type(m) is nn.ReLU:
        return 0.0
    if type(m) is nn.LeakyReLU:
        return float(m.negative_slope)
    return None


def fuse_norm_act(model: nn.Module) -> nn.Module:
    n_bn = n_up = n_pool = 0
    for mod in model.modules():
        if isinstance(mod, nn.Sequential):
            items = list(mod._modules.items())
            for i, (name, m) in enumerate(items):
                if type(m) in (nn.BatchNorm2d, nn.BatchNorm3d) and i > 0 and hasattr(items[i - 1][1], 'local_update'):
                    slope = _slope_of(items[i + 1][1]) if i + 1 < len(items) else None
                    m.__class__ = FusedBatchNormAct2d if type(m) is nn.BatchNorm2d else FusedBatchNormAct3d
                    m._act_slope = 1.0 if slope is None else slope
                    if slope is not None:
                        mod._modules[items[i + 1][0]] = nn.Identity()     # activation now lives in the fused module
                    n_bn += 1
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
    model._hebb_fused = dict(bn_act=n_bn, upsample=n_up, maxpool=n_pool)
    return model
