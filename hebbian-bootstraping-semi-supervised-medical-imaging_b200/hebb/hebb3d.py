"""3-D Hebbian layers — drop-in for the reference module of the same name (hebb/hebb3d.py)."""
import torch.nn as nn

from ._core import _HebbianConvNd, normalize  # noqa: F401

# The reference chunks input channels by this many to bound its materialised unfold
# (hebb3d.py:7,117-125).  Nothing is materialised here; the constant is kept for importers.
PARALLEL_CHANNELS = 32

__all__ = ['PARALLEL_CHANNELS', 'HebbianConv3d', 'HebbianConvTranspose3d']


class HebbianConv3d(_HebbianConvNd):
    """A 3d convolutional layer that learns through Hebbian plasticity (reference hebb3d.py:15-216)."""

    _nd = 3
    _transposed = False

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True,
                 w_nrm=True, act=nn.Identity(),
                 mode=_HebbianConvNd.MODE_SWTA, k=1, patchwise=True,
                 contrast=1., uniformity=False, alpha=0.):
        super().__init__()
        self._setup(in_channels, out_channels, kernel_size, stride, padding, bias, w_nrm, act,
                    mode, k, patchwise, contrast, uniformity, alpha)


class HebbianConvTranspose3d(HebbianConv3d):
    """Transposed twin (reference hebb3d.py:219-305); weight/delta_w are (Cin, Cout, kd, kh, kw) views."""

    MODE_SWTA_T = 'swta_t'
    MODE_HPCA_T = 'hpca_t'
    _transposed = True

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, w_nrm=True,
                 act=nn.Identity(), mode=MODE_SWTA_T, k=1, patchwise=True, contrast=1., uniformity=False,
                 alpha=0.):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, bias, w_nrm, act, mode, k,
                         patchwise, contrast, uniformity, alpha)
