"""2-D Hebbian layers — drop-in for the reference module of the same name (hebb/hebb.py).

Same constructor signatures, attributes, methods and state_dict keys; the arithmetic runs in
libhebb_sm100.so (see _core.py).  CUDA (sm_100) tensors only.
"""
import torch.nn as nn

from ._core import _HebbianConvNd, normalize  # noqa: F401  (normalize is part of the module API)

ADA_STEP = False   # kept for API compatibility; the adaptive step of hebb.py:108-111 is dead code upstream

__all__ = ['ADA_STEP', 'normalize', 'HebbianConv2d', 'HebbianConvTranspose2d']


class HebbianConv2d(_HebbianConvNd):
    """A 2d convolutional layer that learns through Hebbian plasticity (reference hebb.py:16-192)."""

    _nd = 2
    _transposed = False

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True,
                 w_nrm=True, act=nn.Identity(),
                 mode=_HebbianConvNd.MODE_SWTA, k=1, patchwise=True,
                 contrast=1., uniformity=False, alpha=0.):
        super().__init__()
        self._setup(in_channels, out_channels, kernel_size, stride, padding, bias, w_nrm, act,
                    mode, k, patchwise, contrast, uniformity, alpha)


class HebbianConvTranspose2d(HebbianConv2d):
    """Transposed twin (reference hebb.py:195-277); weight/delta_w are (Cin, Cout, kh, kw) views."""

    MODE_SWTA_T = 'swta_t'
    MODE_HPCA_T = 'hpca_t'
    _transposed = True

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, w_nrm=True,
                 act=nn.Identity(), mode=MODE_SWTA_T, k=1, patchwise=True, contrast=1., uniformity=False,
                 alpha=0.):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, bias, w_nrm, act, mode, k,
                         patchwise, contrast, uniformity, alpha)
