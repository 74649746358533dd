"""CPU oracle for the Hebbian convolution hot path — TEST INFRASTRUCTURE ONLY.

This file restates, as plain functional PyTorch-on-CPU arithmetic, what the
reference computes on its Hebbian-pretraining path.  It exists so that the CUDA
product path can be checked against an independent statement of the algorithm
on a box where ``/root/reference`` is not mounted.  Nothing under
``hebbian-bootstraping-semi-supervised-medical-imaging_b200/`` may import it;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do.

Pinning: the reference's own tests hold no golden vectors, so
the oracle is pinned against outputs of the *live reference modules* generated
in the build container by ``tests/golden/make_golden.py`` and committed as
``tests/golden/hebb_golden.npz`` (see ``tests/test_oracle_golden.py``).

Every function cites the reference lines it follows (paths relative to the
reference checkout).  The code is dtype-agnostic: pass float64 tensors to get
the fp64 error-budget variant.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------
# a1  normalize()                       hebb/hebb.py:10-13, hebb/hebb3d.py:9-12
# --------------------------------------------------------------------------
def unit_rows(w: torch.Tensor) -> torch.Tensor:
    """Divide every leading-index slice of ``w`` by its L2 norm (0-norm -> 1)."""
    flat = w.reshape(w.shape[0], -1)
    nrm = flat.pow(2).sum(dim=1, keepdim=True).sqrt()
    nrm = torch.where(nrm == 0, torch.ones_like(nrm), nrm)
    return (flat / nrm).reshape(w.shape)


# --------------------------------------------------------------------------
# a2  pad()                              hebb/hebb.py:83-85, hebb/hebb3d.py:82-84
# --------------------------------------------------------------------------
def pad_list(padding, nd: int):
    """The F.pad argument the reference builds.

    An int pads every side.  A tuple (p0, p1[, p2]) becomes
    [p0,p0,p1,p1(,p2,p2)], which F.pad applies starting from the LAST dim, so
    p0 lands on W and p1 on H — swapped w.r.t. nn.Conv2d.  Kept on purpose.
    """
    if isinstance(padding, int):
        return [padding] * (2 * nd)
    padding = list(padding)
    if len(padding) == nd:
        out = []
        for p in padding:
            out += [p, p]
        return out
    return padding


def zero_halo(x: torch.Tensor, padding, nd: int) -> torch.Tensor:
    return F.pad(x, pad_list(padding, nd))


# --------------------------------------------------------------------------
# a3  compute_activation()               hebb/hebb.py:70-81, hebb/hebb3d.py:69-80
# --------------------------------------------------------------------------
def conv_activation(xp, weight, bias, stride, w_nrm=True):
    """y = conv(x_padded, W/|W|, b, stride); act is Identity on this path."""
    w = unit_rows(weight) if w_nrm else weight
    nd = xp.dim() - 2
    conv = F.conv2d if nd == 2 else F.conv3d
    return conv(xp, w, bias=bias, stride=stride)


def convT_activation(x, weight, bias, stride, w_nrm=True):
    """Transposed twin (hebb.py:226-232, hebb3d.py:250-256).

    ``weight`` is the (Cin, Cout, k...) view; normalisation runs over every dim
    but the first, i.e. per INPUT channel (hebb3d.py:78 with the transposed view).
    """
    w = unit_rows(weight) if w_nrm else weight
    nd = x.dim() - 2
    convT = F.conv_transpose2d if nd == 2 else F.conv_transpose3d
    return convT(x, w, bias=bias, stride=stride)


# --------------------------------------------------------------------------
# a4  patch matrix                       hebb/hebb.py:105-106, hebb/hebb3d.py:92-101
# --------------------------------------------------------------------------
def patch_matrix(xp, kernel, stride):
    """X[P, K]: row order (b, out-spatial...), column order (c, k-spatial...).

    Written with Tensor.unfold on every spatial dim rather than F.unfold so it
    is one code path for 2-D and 3-D.
    """
    nd = xp.dim() - 2
    t = xp
    for d in range(nd):
        t = t.unfold(2 + d, kernel[d], stride[d])
    # t: (B, C, o1..ond, k1..knd)
    perm = [0] + list(range(2, 2 + nd)) + [1] + list(range(2 + nd, 2 + 2 * nd))
    t = t.permute(*perm)
    P = math.prod(t.shape[: 1 + nd])
    return t.reshape(P, -1)


# --------------------------------------------------------------------------
# a5  soft winner-take-all               hebb/hebb.py:107, hebb/hebb3d.py:116
# --------------------------------------------------------------------------
def swta_response(y, k):
    """r[Cout, P] = softmax over channels of k*y, pixels in (b, spatial) order."""
    r = (y * k).softmax(dim=1)
    return r.transpose(0, 1).reshape(y.shape[1], -1)


def winners(y):
    """Hard-WTA index per pixel: argmax over channels (first max wins)."""
    return y.argmax(dim=1)


# --------------------------------------------------------------------------
# a6+a7  SWTA weight delta (patchwise)   hebb/hebb.py:112-115, hebb/hebb3d.py:115-125
# --------------------------------------------------------------------------
def swta_delta(xp, y, weight, k, stride):
    """dW_c = sum_p r[c,p] X[p,:] - (sum_p r[c,p]) W_c   (un-normalised W)."""
    kernel = tuple(weight.shape[2:])
    X = patch_matrix(xp, kernel, stride)
    r = swta_response(y, k)
    w2 = weight.reshape(weight.shape[0], -1)
    return (r @ X - r.sum(dim=1, keepdim=True) * w2).reshape(weight.shape)


# --------------------------------------------------------------------------
# 8f-1  HPCA / Sanger delta (patchwise)  hebb/hebb.py:122-135, hebb/hebb3d.py:139-153
# --------------------------------------------------------------------------
def hpca_delta(xp, y, weight, stride):
    """dW = y X - tril(y y^T) W : the response is the layer output itself, the decay couples
    filter c to every filter c' <= c (generalised Hebbian algorithm)."""
    kernel = tuple(weight.shape[2:])
    X = patch_matrix(xp, kernel, stride)
    r = y.transpose(0, 1).reshape(y.shape[1], -1)
    C = weight.shape[0]
    low = torch.tril(torch.ones(C, C, dtype=weight.dtype))
    w2 = weight.reshape(C, -1)
    return (r @ X - ((r @ r.t()) * low) @ w2).reshape(weight.shape)


def hpca_exchanged_delta(x, y, weight, stride):
    """Transposed layer in mode 'hpca' (hebb.py:243-246): the plain-conv rule with the roles of x and y
    swapped — response = the layer INPUT x, patches = unfold(OUTPUT y); weight is the (Cin, Cout, k...) view."""
    kernel = tuple(weight.shape[2:])
    Y = patch_matrix(y, kernel, stride)                  # (P, Cout*taps)
    r = x.transpose(0, 1).reshape(x.shape[1], -1)        # (Cin, P)
    C = weight.shape[0]
    low = torch.tril(torch.ones(C, C, dtype=weight.dtype))
    w2 = weight.reshape(C, -1)
    return (r @ Y - ((r @ r.t()) * low) @ w2).reshape(weight.shape)


# --------------------------------------------------------------------------
# a9  transposed SWTA delta              hebb/hebb.py:252-264, hebb/hebb3d.py:276-289
# --------------------------------------------------------------------------
def swta_t_delta(x, y, weight, k, stride):
    """delta for the (Cin, Cout, k...) weight of a transposed conv.

    r = softmax_c(k*y) on the up-sampled output, cut into kernel-size patches
    at the layer's stride (the reference unfolds r, hebb.py:256); Hebbian term
    H[ci,co,off] = sum_p r[co,off,p] x[ci,p]; decay, with patchwise=True, is
    sum over ALL offsets of (sum_p r[co,off',p]) W[ci,co,off'] broadcast back to
    every offset (hebb.py:262-263).
    Requires the patch grid of r to have as many patches as x has pixels,
    which holds for kernel == stride (the only shape the networks use).
    """
    nd = x.dim() - 2
    kernel = tuple(weight.shape[2:])
    Cin, Cout = weight.shape[0], weight.shape[1]
    r = (y * k).softmax(dim=1)
    R = patch_matrix(r, kernel, stride)              # (P, Cout*prod(kernel)), cols (co, off)
    P = R.shape[0]
    nk = math.prod(kernel)
    R = R.reshape(P, Cout, nk)
    perm = [0] + list(range(2, 2 + nd)) + [1]
    xf = x.permute(*perm).reshape(-1, Cin)           # (P, Cin)
    H = torch.einsum('pco,pi->ico', R, xf)           # (Cin, Cout, nk)
    rs = R.sum(dim=0)                                # (Cout, nk)
    w3 = weight.reshape(Cin, Cout, nk)
    dec = (rs.unsqueeze(0) * w3).sum(dim=2, keepdim=True)   # (Cin, Cout, 1)
    return (H - dec).reshape(weight.shape)


# --------------------------------------------------------------------------
# §8f-4  contrastive rule                hebb/hebb.py:143-172, hebb/hebb3d.py:167-197
# --------------------------------------------------------------------------
def _box_sum(t):
    """Sum over the 3^nd neighbourhood (zero beyond the border): unfold(.., 3, padding=1).sum(-1) of the reference."""
    nd = t.dim() - 2
    pool = F.avg_pool2d if nd == 2 else F.avg_pool3d
    return pool(t, 3, stride=1, padding=1, count_include_pad=True) * float(3 ** nd)


def _unit_channels(t):
    nrm = t.pow(2).sum(dim=1, keepdim=True).sqrt()
    return t / torch.where(nrm == 0, torch.ones_like(nrm), nrm)


def contrastive_delta(xp, weight, bias, stride, contrast=1., perm=None, w_nrm=True):
    """delta_w (and dL/dbias) of the reference's contrastive rule for a plain conv layer on the padded input xp:
    L = sum_pixels [ -(S*y) + contrast * (S[perm]*y) ],  y = channel-normalised layer output, S = 3^nd box sum of
    y, perm = a permutation of the batch (the reference draws torch.randperm(B) at this point).
    delta_w += dL/dW; dL/dbias lands in bias.grad as a side effect of the reference's L.backward().
    (uniformity=True is not restated: that branch of the reference raises for Cout > 1 -- apply_weights() adds
    the [Cout] bias to a 1-channel map, hebb.py:75,160.)"""
    w = weight.detach().clone().requires_grad_(True)
    b = bias.detach().clone().requires_grad_(True) if bias is not None else None
    y = _unit_channels(conv_activation(xp.detach(), w, b, stride, w_nrm))
    S = _box_sum(y)
    if perm is None:
        perm = torch.randperm(y.shape[0])
    L = (-(S * y) + contrast * S[perm] * y).sum()
    grads = torch.autograd.grad(L, [w] + ([b] if b is not None else []), allow_unused=True)
    return grads[0], (grads[1] if b is not None else None)


def hpca_t_delta(x, y, weight, stride, patchwise=True, nd=None):
    """mode 'hpca_t' of the transposed layers (hebb/hebb.py:266-277, hebb/hebb3d.py:291-305), kernel == stride, no
    padding.  weight: the (Cin, Cout, k...) view.  For each kernel offset t the responses are y at offset t of every
    input position's output patch; delta[ci, co, t] = sum_p y_t[co,p] x[ci,p] - sum_{co'<=co} (y_t y_t^T)[co,co'] W[ci,co',t]
    with the decay summed over t when patchwise.  The 3-D reference applies the triangular mask inside chunks of 32
    output channels (PARALLEL_CHANNELS, hebb3d.py:12)."""
    nd = x.dim() - 2 if nd is None else nd
    ks = tuple(weight.shape[2:])
    B, C = y.shape[0], y.shape[1]
    taps = 1
    for k in ks:
        taps *= k
    shape = [B, C]
    for n, k in zip(x.shape[2:], ks):
        shape += [n, k]
    k_axes = [3 + 2 * i for i in range(nd)]
    s_axes = [2 + 2 * i for i in range(nd)]
    r = y.reshape(shape).permute(*k_axes, 1, 0, *s_axes).reshape(taps, C, -1)
    xf = x.permute(0, *range(2, nd + 2), 1).reshape(-1, x.shape[1])
    wp = weight.permute(*range(2, nd + 2), 1, 0).reshape(taps, C, -1)
    out = torch.zeros(weight.shape, dtype=weight.dtype)
    step = 32 if nd == 3 else C
    for c0 in range(0, C, step):
        c1 = min(C, c0 + step)
        ri = r[:, c0:c1]
        tri = torch.tril(torch.ones(c1 - c0, c1 - c0, dtype=x.dtype))
        dec = (ri.matmul(ri.transpose(-2, -1)) * tri).matmul(wp[:, c0:c1])
        if patchwise:
            dec = dec.sum(dim=0, keepdim=True)
        out[:, c0:c1] = (ri.matmul(xf.unsqueeze(0)) - dec).permute(2, 1, 0).reshape(out[:, c0:c1].shape)
    return out


# --------------------------------------------------------------------------
# a8  local_update()                     hebb/hebb.py:174-192, hebb/hebb3d.py:198-216
# --------------------------------------------------------------------------
def fold_delta_into_grad(grad: Optional[torch.Tensor], delta_w: torch.Tensor, alpha: float):
    """Returns (new_grad, zeroed delta_w)."""
    if grad is None:
        new = -alpha * delta_w
    else:
        new = (1 - alpha) * grad - alpha * delta_w
    return new, torch.zeros_like(delta_w)


# --------------------------------------------------------------------------
# Module wrappers so whole networks can be run through the oracle
# (needed for the C2/C4 workloads and the CPU-baseline timing).
# --------------------------------------------------------------------------
def _tuple(v, nd):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * nd


class OracleHebbConv(nn.Module):
    """HebbianConv{2,3}d in SWTA/patchwise mode (hebb.py:16-192, hebb3d.py:15-216)."""

    def __init__(self, nd, cin, cout, kernel, stride=1, padding=0, bias=True,
                 w_nrm=True, k=1., alpha=0., act=None):
        super().__init__()
        self.nd = nd
        self.kernel_size = _tuple(kernel, nd)
        self.stride = _tuple(stride, nd)
        self.padding = padding
        self.weight = nn.Parameter(torch.empty(cout, cin, *self.kernel_size))
        nn.init.xavier_normal_(self.weight)
        self.bias = nn.Parameter(torch.zeros(cout), requires_grad=bias)
        self.register_buffer('delta_w', torch.zeros_like(self.weight))
        self.w_nrm, self.k, self.alpha = w_nrm, k, alpha
        self.act = act if act is not None else nn.Identity()

    def forward(self, x):
        xp = zero_halo(x, self.padding, self.nd)
        # the rule sees the ACTIVATED output: y = act(conv(...)) feeds compute_update (hebb.py:80,87-90,107)
        y = self.act(conv_activation(xp, self.weight, self.bias, self.stride, self.w_nrm))
        if self.training and self.alpha != 0:
            with torch.no_grad():
                self.delta_w += swta_delta(xp, y, self.weight, self.k, self.stride)
        return y

    @torch.no_grad()
    def local_update(self):
        self.weight.grad, z = fold_delta_into_grad(self.weight.grad, self.delta_w, self.alpha)
        self.delta_w.zero_()


class OracleHebbConvT(nn.Module):
    """HebbianConvTranspose{2,3}d in swta_t mode (hebb.py:195-264, hebb3d.py:219-289)."""

    def __init__(self, nd, cin, cout, kernel, stride=1, padding=0, bias=True,
                 w_nrm=True, k=1., alpha=0.):
        super().__init__()
        self.nd = nd
        self.kernel_size = _tuple(kernel, nd)
        self.stride = _tuple(stride, nd)
        self.padding = padding
        w = torch.empty(cout, cin, *self.kernel_size)
        nn.init.xavier_normal_(w)
        # stored as the (Cin, Cout, ...) transposed VIEW, like the reference
        self.weight = nn.Parameter(w)
        with torch.no_grad():
            self.weight.transpose_(0, 1)
        self.bias = nn.Parameter(torch.zeros(cout), requires_grad=bias)
        self.register_buffer('delta_w', torch.zeros(cout, cin, *self.kernel_size))
        with torch.no_grad():
            self.delta_w.transpose_(0, 1)
        self.w_nrm, self.k, self.alpha = w_nrm, k, alpha

    def forward(self, x):
        xp = zero_halo(x, self.padding, self.nd)
        y = convT_activation(xp, self.weight, self.bias, self.stride, self.w_nrm)
        if self.training and self.alpha != 0:
            with torch.no_grad():
                self.delta_w += swta_t_delta(xp, y, self.weight, self.k, self.stride)
        return y

    @torch.no_grad()
    def local_update(self):
        self.weight.grad, z = fold_delta_into_grad(self.weight.grad, self.delta_w, self.alpha)
        self.delta_w.zero_()


def oracle_makehebbian(model: nn.Module, exclude: Optional[Sequence[str]] = None,
                       k: float = 50., alpha: float = 1., w_nrm: bool = True) -> nn.Module:
    """Module surgery with the reference's semantics (hebb/makehebbian.py:45-87),
    restricted to the layer kinds the benchmark networks contain."""
    exclude = list(exclude or [])
    roots = [m for n, m in model.named_modules() if n in exclude]
    skipped = {id(s) for r in roots for s in r.modules()}

    def visit(parent):
        for name, child in list(parent.named_children()):
            if id(child) in skipped:
                continue
            t = type(child)
            if t in (nn.Conv2d, nn.Conv3d):
                nd = 2 if t is nn.Conv2d else 3
                new = OracleHebbConv(nd, child.in_channels, child.out_channels, child.kernel_size,
                                     child.stride, child.padding, False, w_nrm, k, alpha)
                nn.init.kaiming_normal_(new.weight.data, a=0, mode='fan_in')
                parent.register_module(name, new)
            elif t in (nn.ConvTranspose2d, nn.ConvTranspose3d):
                nd = 2 if t is nn.ConvTranspose2d else 3
                new = OracleHebbConvT(nd, child.in_channels, child.out_channels, child.kernel_size,
                                      child.stride, child.padding, False, w_nrm, k, alpha)
                nn.init.kaiming_normal_(new.weight.data, a=0, mode='fan_in')
                parent.register_module(name, new)
            else:
                for p in child.parameters(recurse=False):
                    p.requires_grad = False

    model.apply(visit)
    return model
