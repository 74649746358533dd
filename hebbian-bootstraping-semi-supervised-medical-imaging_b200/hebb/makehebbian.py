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


def _hebbian_twin(child, hebb_params):
    """The Hebbian replacement for one stock layer, or None if `child` is not convertible."""
    kind = type(child)
    if kind in _CONV_TWINS:
        twin, plain_rule, unit = _CONV_TWINS[kind]
        if child.dilation != 1 and child.dilation != unit:
            raise RuntimeError("Dilation not supported with Hebbian layers")
        if child.groups != 1:
            raise RuntimeError("Grouped convolution not supported with Hebbian layers")
        kwargs = adjust_hebbian_params(hebb_params) if plain_rule else hebb_params
        layer = twin(child.in_channels, child.out_channels, child.kernel_size, child.stride, child.padding, False, **kwargs)
        return init_weights(layer, init_type='kaiming')
    if kind is nn.Linear:
        # a fully connected layer becomes a 1x1 Hebbian conv between two reshape shims (bias stays trainable)
        conv = HebbianConv2d(child.in_features, child.out_features, 1, 1, **adjust_hebbian_params(hebb_params))
        return nn.Sequential(UnsqueezeLast(2), init_weights(conv, init_type='kaiming'), FlattenLast(2))
    return None


def makehebbian(model, exclude=None, hebb_params=None):
    """In-place conversion; returns `model`.  `exclude` holds exact `named_modules()` names whose whole
    sub-trees stay as they are (and stay trainable by back-prop)."""
    hebb_params = default_hebb_params if hebb_params is None else hebb_params
    names = set(exclude or ())
    roots = [(n, m) for n, m in model.named_modules() if n in names]
    print("Layers excluded from conversion to Hebbian: {}".format([n for n, _ in roots]))
    protected = {id(sub) for _, root in roots for sub in root.modules()}

    # snapshot of the ORIGINAL tree: layers created below are never revisited (their parameters stay trainable)
    edges = [(parent, name, child) for parent in list(model.modules()) for name, child in list(parent.named_children())]
    for parent, name, child in edges:
        if id(child) in protected:
            continue
        new = _hebbian_twin(child, hebb_params)
        if new is not None:
            parent.register_module(name, new)
        if new is None or type(child) in (nn.Conv2d, nn.ConvTranspose2d):
            # everything that is not converted is frozen; so are the orphaned parameters of a replaced
            # 2-D conv (reference quirk: makehebbian.py:64 and :72 are separate if-chains)
            for prm in child.parameters(recurse=False):
                prm.requires_grad = False
    return model
