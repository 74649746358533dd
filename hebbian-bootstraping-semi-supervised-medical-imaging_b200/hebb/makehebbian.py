"""makehebbian(): swap every Conv/ConvTranspose/Linear of a model for its Hebbian twin.

Behavioural twin of the reference hebb/makehebbian.py:45-87 (exact-name exclusion that covers
whole sub-trees, kaiming init of the new layers, Linear -> 1x1 conv between reshape shims,
requires_grad=False on every other directly-owned parameter, RuntimeError on dilation/groups).
"""
import torch.nn as nn

from .hebb import HebbianConv2d, HebbianConvTranspose2d
from .hebb3d import HebbianConv3d, HebbianConvTranspose3d

default_hebb_params = {'w_nrm': True, 'act': nn.Identity(), 'mode': HebbianConvTranspose2d.MODE_SWTA_T, 'k': 50,
                       'patchwise': True, 'contrast': 1., 'uniformity': False, 'alpha': 0.}


class UnsqueezeLast(nn.Module):
    """Append d singleton dims (lets a Linear run as a 1x1 convolution)."""

    def __init__(self, d=2):
        super().__init__()
        self.d = d

    def forward(self, x):
        return x.reshape(*x.shape, *([1] * self.d))


class FlattenLast(nn.Module):
    """Merge the last d+1 dims into one."""

    def __init__(self, d=2):
        super().__init__()
        self.d = d

    def forward(self, x):
        return x.reshape(*(x.shape[:-self.d - 1]), -1)


def adjust_hebbian_params(hebb_params):
    """Non-transposed layers take the plain rule: 'swta_t' -> 'swta', 'hpca_t' -> 'hpca'."""
    out = hebb_params.copy()
    mode = out.get('mode', None)
    if mode is not None and mode.endswith('_t'):
        out['mode'] = mode[:-2]
    return out


_INITS = {
    'normal': lambda w, gain: nn.init.normal_(w, 0.0, gain),
    'xavier': lambda w, gain: nn.init.xavier_normal_(w, gain=gain),
    'kaiming': lambda w, gain: nn.init.kaiming_normal_(w, a=0, mode='fan_in'),
    'orthogonal': lambda w, gain: nn.init.orthogonal_(w, gain=gain),
}


def init_weights(m, init_type='normal', gain=0.02):
    if init_type not in _INITS:
        raise NotImplementedError("Unsupported initialization method {}".format(init_type))
    _INITS[init_type](m.weight.data, gain)
    return m


_CONV_TWINS = {
    nn.Conv2d: (HebbianConv2d, True, (1, 1)),
    nn.ConvTranspose2d: (HebbianConvTranspose2d, False, (1, 1)),
    nn.Conv3d: (HebbianConv3d, True, (1, 1, 1)),
    nn.ConvTranspose3d: (HebbianConvTranspose3d, False, (1, 1, 1)),
}


def makehebbian(model, exclude=None, hebb_params=None):
    if hebb_params is None:
        hebb_params = default_hebb_params
    wanted = list(exclude) if exclude is not None else []
    roots = [(n, m) for n, m in model.named_modules() if n in wanted]
    print("Layers excluded from conversion to Hebbian: {}".format([n for n, _ in roots]))
    kept = [s for _, r in roots for s in r.modules()]

    def convert(parent):
        for name, child in parent.named_children():
            if any(child is k for k in kept):
                continue
            kind = type(child)
            if kind in _CONV_TWINS:
                twin, adjust, unit = _CONV_TWINS[kind]
                if child.dilation != 1 and child.dilation != unit:
                    raise RuntimeError("Dilation not supported with Hebbian layers")
                if child.groups != 1:
                    raise RuntimeError("Grouped convolution not supported with Hebbian layers")
                params = adjust_hebbian_params(hebb_params) if adjust else hebb_params
                new = twin(child.in_channels, child.out_channels, child.kernel_size, child.stride, child.padding,
                           False, **params)
                parent.register_module(name, init_weights(new, init_type='kaiming'))
                if kind in (nn.Conv2d, nn.ConvTranspose2d):
                    # reference quirk (makehebbian.py:64 vs :72 are separate if-chains): the replaced
                    # 2-D layer also reaches the final else and has its own parameters frozen
                    for p in child.parameters(recurse=False):
                        p.requires_grad = False
            elif kind is nn.Linear:
                conv = HebbianConv2d(child.in_features, child.out_features, 1, 1, **adjust_hebbian_params(hebb_params))
                parent.register_module(name, nn.Sequential(UnsqueezeLast(2), init_weights(conv, init_type='kaiming'),
                                                           FlattenLast(2)))
            else:
                for p in child.parameters(recurse=False):
                    p.requires_grad = False

    model.apply(convert)
    return model
