"""CPU-side tests: the C-ABI library loads and exports what include/hebb_sm100.h declares,
makehebbian() reproduces the reference's module surgery (fixtures from the live reference),
the restated workload networks match the reference's, and the drop-in refuses CPU tensors."""
import ctypes
import io
import contextlib
import json
import os
import re

import pytest
import torch
import torch.nn as nn

import hebb
from hebb import _native
from hebb.makehebbian import (makehebbian, UnsqueezeLast, FlattenLast, adjust_hebbian_params,
                              default_hebb_params, init_weights)
import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'hebb_sm100.h')).read()
    declared = set(re.findall(r'\b(hebb_[a-z_A-Z0-9]+)\s*\(', header))
    declared -= {'hebb_status', 'hebb_prec'}
    assert {'hebb_conv_swta_step', 'hebb_convT_swta_step', 'hebb_local_update_multi', 'hebb_wnorm'} <= declared
    lib = ctypes.CDLL(_native.lib_path())
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in hebb_sm100.h but not exported'
    assert set(_native.EXPORTS) <= declared
    assert b'sm_100' in _native.load().hebb_version()
    assert _native.load().hebb_status_str(-1).decode().startswith('no sm_100')


def test_geometry_queries_need_no_gpu():
    d = _native.make_desc(2, 8, 3, 64, (128, 128), (3, 3), (1, 1), (1, 1), (1, 1), False)
    assert _native.out_shape(d) == [128, 128]
    assert _native.workspace_bytes(d, _native.PREC_FP32) > 0
    assert _native.workspace_bytes(d, _native.PREC_BF16X3) > 0
    d3 = _native.make_desc(3, 2, 8, 4, (3, 4, 5), (2, 2, 2), (2, 2, 2), (0, 0, 0), (0, 0, 0), True)
    assert _native.out_shape(d3) == [6, 8, 10]
    bad = _native.make_desc(2, 1, 3, 8, (2, 2), (3, 3), (1, 1), (0, 0), (0, 0), False)
    with pytest.raises(RuntimeError):
        _native.out_shape(bad)


def _desc3(B, Cin, Cout, sp, k=3, transposed=False):
    nd = len(sp)
    return _native.make_desc(nd, B, Cin, Cout, sp, (k,) * nd, ((k if transposed else 1),) * nd,
                             ((0 if transposed else k // 2),) * nd, ((0 if transposed else k // 2),) * nd, transposed)


def test_planner_keeps_single_wave_and_tiles_every_c4_layer():
    """Host-side planner (no GPU needed).  147 position tiles of the 12x12x10 layers must stay ONE wave on
    148 SMs (a padded position count once pushed them to 152 = two waves = 2x the time), and every layer
    of the two benchmark networks must get a tensor-core plan."""
    for cin, cout in ((256, 512), (512, 512), (1024, 512)):
        p = _native.plan(_desc3(8, cin, cout, (12, 12, 10)), _native.PREC_BF16X3)
        assert p is not None and p['f_tiles'] == 147 and p['f_tiles'] * p['n_ct'] <= 148
    for cin, cout, sp in ((1, 64, (96, 96, 80)), (64, 64, (96, 96, 80)), (128, 64, (96, 96, 80)), (128, 128, (48, 48, 40)),
                          (256, 256, (24, 24, 20)), (512, 1024, (6, 6, 5)), (1024, 1024, (6, 6, 5))):
        for prec in (_native.PREC_BF16X3, _native.PREC_BF16):
            p = _native.plan(_desc3(8, cin, cout, sp), prec)
            assert p is not None, (cin, cout, sp)
            assert p['f_smem'] <= 227 * 1024 and p['d_smem'] <= 227 * 1024 and p['f_tmem'] <= 512 and p['d_tmem'] <= 512
            # the update kernel is launched as one wave
            assert p['ngrp'] * p['n_cin'] * p['n_cout'] * p['PS'] <= 148 or p['PS'] == 1
    for cin, cout, sp in ((3, 16, (256, 256)), (16, 16, (256, 256)), (32, 16, (256, 256)), (256, 128, (16, 16))):
        assert _native.plan(_desc3(64, cin, cout, sp), _native.PREC_BF16X3) is not None
    for cin, cout, sp in ((1024, 512, (6, 6, 5)), (256, 128, (24, 24, 20)), (128, 64, (48, 48, 40))):
        assert _native.plan(_desc3(8, cin, cout, sp, k=2, transposed=True), _native.PREC_BF16X3) is not None


def test_planner_collector_reuse_cases_are_the_ones_the_gpu_tests_run():
    """The contraction kernel has two loop orders (tap outer / k-step outer with A-operand collector re-use).  The
    GPU parity cases named *_run6 / *_run9 exist to exercise the second one in both precision modes: pin that the
    planner really picks it for them, and that the full-size 64 -> 64 layer of the 3-D network uses it."""
    for cin, cout, sp in ((64, 64, (16, 16, 16)), (128, 32, (10, 10, 10))):
        for prec in (_native.PREC_BF16X3, _native.PREC_BF16):
            p = _native.plan(_desc3(1, cin, cout, sp), prec)
            assert p is not None and p['reuse'] == 1, (cin, cout, prec, p)
    p = _native.plan(_desc3(8, 64, 64, (96, 96, 80)), _native.PREC_BF16)
    assert p['reuse'] == 1 and p['nrep'] == 2 and p['by_kh'] == 0
    # wide response tiles are math-bound: no re-use
    assert _native.plan(_desc3(8, 128, 128, (48, 48, 40)), _native.PREC_BF16)['reuse'] == 0


def test_planner_swizzled_response_variant_only_where_it_applies():
    """64 response channels (128 with 64 input channels), 3-wide kernel rows, 64 or k*128 input channels: the update
    reads the responses as SWIZZLE_128B [position][64] images (N = 192 instructions, one kernel row each), in both
    tensor-core precision modes.  Everything else keeps the plain planes."""
    for cin, sp in ((64, (96, 96, 80)), (128, (96, 96, 80)), (64, (8, 8, 8)), (256, (12, 12, 10))):
        for prec in (_native.PREC_BF16, _native.PREC_BF16X3):
            p = _native.plan(_desc3(8, cin, 64, sp), prec)
            assert p['rsw'] == 1 and p['ws_MiB'] >= 0, (cin, sp, prec)
    assert _native.plan(_desc3(64, 128, 64, (64, 64)), _native.PREC_BF16)['rsw'] == 1          # 2-D
    assert _native.plan(_desc3(8, 64, 128, (48, 48, 40)), _native.PREC_BF16)['rsw'] == 1       # two 64-channel response planes
    for cin, cout in ((32, 64), (128, 128), (256, 128), (64, 32), (96, 64)):
        assert _native.plan(_desc3(8, cin, cout, (24, 24, 20)), _native.PREC_BF16)['rsw'] == 0, (cin, cout)
    assert _native.plan(_desc3(8, 64, 64, (24, 24, 20), k=1), _native.PREC_BF16)['rsw'] == 0


def test_planner_invariants_over_a_shape_grid():
    """Every plan the tensor-core planner returns must fit the machine: shared memory <= 227 KB, TMEM <= 512 columns,
    the update kernel one wave (or unsplit), the swizzled-response variant 1024-byte-aligned stages within the same
    limits.  Swept over channel counts, image sizes, 2-D/3-D and both tensor-core precision modes."""
    n = 0
    for nd, sps in ((2, ((16, 16), (64, 64), (256, 256), (33, 17))), (3, ((6, 6, 5), (12, 12, 10), (48, 48, 40), (96, 96, 80)))):
        for sp in sps:
            for cin in (1, 3, 16, 32, 64, 128, 256, 512):
                for cout in (16, 32, 64, 128, 256, 1024):
                    if nd == 3 and max(sp) >= 48 and cin * cout > 128 * 128:
                        continue                                  # beyond the networks' shapes (and the workspace)
                    for prec in (_native.PREC_BF16X3, _native.PREC_BF16):
                        p = _native.plan(_desc3(4, cin, cout, sp), prec)
                        if p is None:
                            continue
                        n += 1
                        assert p['f_smem'] <= 227 * 1024 and p['d_smem'] <= 227 * 1024, (nd, sp, cin, cout, prec, p)
                        assert p['f_tmem'] <= 512 and p['d_tmem'] <= 512
                        assert p['ngrp'] * p['n_cin'] * p['n_cout'] * p['PS'] <= 148 or p['PS'] == 1
                        if p['rsw']:
                            assert cout in (64, 128) and (cin == 64 or cin % 128 == 0)
                            assert p['rs_smem'] <= 226 * 1024 and p['rs_tmem'] <= 512 and p['rs_BLK'] % 16 == 0
                            assert p['rs_PS'] >= 1 and p['rs_stackM'] == (1 if (cin == 64 and prec == _native.PREC_BF16X3) else 0)
    assert n > 500


def test_no_cpu_fallback():
    layer = hebb.HebbianConv2d(3, 8, 3, padding=1, alpha=1.)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        layer(torch.randn(1, 3, 8, 8))
    with pytest.raises(RuntimeError):
        layer.local_update()


class Net(nn.Module):   # the reference test's toy net (tests/test_makehebbian.py:5-39)
    def __init__(self):
        super().__init__()
        self.down = nn.Sequential(nn.Conv2d(3, 16, 3, stride=2), nn.BatchNorm2d(16), nn.ReLU())
        self.up = nn.Sequential(nn.ConvTranspose2d(16, 20, 3, stride=2), nn.BatchNorm2d(20), nn.ReLU())
        self.clf = nn.Sequential(FlattenLast(2), nn.Linear(20, 16), nn.BatchNorm1d(16), nn.ReLU(),
                                 nn.Dropout(0.5), nn.Linear(16, 10))

    def forward(self, x):
        return self.clf(self.up(self.down(x)))


def describe(net):
    mods = {n: type(m).__name__ for n, m in net.named_modules()}
    params = {n: [list(p.shape), bool(p.requires_grad), list(p.stride())] for n, p in net.named_parameters()}
    bufs = {n: list(b.shape) for n, b in net.named_buffers()}
    hp = {n: dict(mode=m.mode, k=m.k, alpha=m.alpha, w_nrm=m.w_nrm, patchwise=m.patchwise,
                  kernel_size=list(m.kernel_size), stride=list(m.stride),
                  padding=(m.padding if isinstance(m.padding, int) else list(m.padding)))
          for n, m in net.named_modules() if hasattr(m, 'local_update')}
    return dict(modules=mods, params=params, buffers=bufs, hebb=hp, state_keys=list(net.state_dict().keys()))


@pytest.mark.parametrize('case,kwargs', [
    ('toy_default_params', dict(exclude=['clf.5'], hebb_params={})),
    ('toy_swta_t', dict(exclude=['clf.5'], hebb_params={'mode': 'swta_t', 'k': 50., 'w_nrm': True, 'alpha': 1.})),
    ('toy_none', dict(exclude=None, hebb_params=None)),
])
def test_makehebbian_matches_reference_structure(golden_meta, case, kwargs):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        net = makehebbian(Net(), **kwargs)
    got = json.loads(json.dumps(describe(net)))
    want = golden_meta['structures'][case]
    for key in ('modules', 'params', 'buffers', 'hebb', 'state_keys'):
        assert got[key] == want[key], key
    assert buf.getvalue().startswith('Layers excluded from conversion to Hebbian:')


def test_makehebbian_helpers(golden_meta):
    want = golden_meta['structures']['adjust']
    assert adjust_hebbian_params({'mode': 'swta_t', 'k': 3}) == want['swta_t']
    assert adjust_hebbian_params({'mode': 'hpca_t'}) == want['hpca_t']
    assert adjust_hebbian_params({'mode': 'swta'}) == want['swta']
    assert adjust_hebbian_params({'k': 2}) == want['none']
    dflt = {k: (v if not isinstance(v, nn.Module) else type(v).__name__) for k, v in default_hebb_params.items()}
    assert dflt == golden_meta['structures']['default_hebb_params']
    with pytest.raises(NotImplementedError):
        init_weights(nn.Conv2d(1, 1, 1), 'bogus')
    with pytest.raises(RuntimeError, match='Dilation'):
        makehebbian(nn.Sequential(nn.Conv2d(1, 1, 3, dilation=2)))
    with pytest.raises(RuntimeError, match='Grouped'):
        makehebbian(nn.Sequential(nn.Conv3d(2, 2, 3, groups=2)))
    x = torch.randn(4, 5)
    assert UnsqueezeLast(2)(x).shape == (4, 5, 1, 1)
    assert FlattenLast(2)(torch.randn(4, 5, 2, 3)).shape == (4, 30)


def test_layer_api_surface():
    l2 = hebb.HebbianConv2d(3, 8, 3, stride=1, padding=1, bias=False, w_nrm=True, mode='swta', k=5, alpha=1.)
    assert list(l2.state_dict().keys()) == ['weight', 'bias', 'delta_w']
    assert l2.weight.shape == (8, 3, 3, 3) and l2.delta_w.shape == (8, 3, 3, 3)
    assert l2.bias.requires_grad is False and l2.kernel_size == (3, 3) and l2.stride == (1, 1)
    for attr in ('mode', 'in_channels', 'out_channels', 'padding', 'w_nrm', 'act', 'k', 'patchwise',
                 'contrast', 'uniformity', 'alpha'):
        assert hasattr(l2, attr)
    assert hebb.HebbianConv2d.MODE_SWTA == 'swta' and hebb.HebbianConvTranspose3d.MODE_SWTA_T == 'swta_t'
    t3 = hebb.HebbianConvTranspose3d(4, 6, 2, stride=2)
    assert t3.mode == 'swta_t' and t3.weight.shape == (4, 6, 2, 2, 2)
    assert t3.weight.stride() == (8, 32, 4, 2, 1) and not t3.weight.is_contiguous()
    l2.mode = 'bogus'
    l2.train()
    with pytest.raises(NotImplementedError, match='Learning mode bogus unavailable'):
        l2._check_mode()
    # padding quirk of hebb.py:84: tuple (p0, p1) pads W with p0 and H with p1
    q = hebb.HebbianConv2d(1, 1, 3, padding=(2, 1))
    assert q._pad_lo_hi() == ([1, 2], [1, 2])
    assert q.pad(torch.zeros(1, 1, 4, 4)).shape == (1, 1, 6, 8)
    # normalize() stays a general utility on CPU tensors
    w = torch.tensor([[3., 4.], [0., 0.]])
    assert torch.allclose(hebb.normalize(w, dim=1), torch.tensor([[0.6, 0.8], [0., 0.]]))


def test_state_dict_round_trip_with_reference_layout():
    """Reference checkpoints store the transposed layers' (Cin, Cout, ...) view contiguously."""
    net = makehebbian(Net(), exclude=['clf.5'], hebb_params={'mode': 'swta_t', 'k': 5., 'alpha': 1.})
    sd = {k: torch.randn_like(v) if v.is_floating_point() else v.clone() for k, v in net.state_dict().items()}
    net.load_state_dict(sd)
    assert torch.equal(net.up[0].weight.detach(), sd['up.0.weight'])
    assert net.up[0].weight.stride() == (9, 144, 3, 1)       # still the transposed view of (Cout,Cin,..)


def test_workload_topologies_match_reference():
    keys = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'workload_state_keys.json')))
    n2 = workloads.unet2d(3, 2)
    n3 = workloads.unet3d(1, 2, init_features=8)
    assert [[n, list(v.shape)] for n, v in n2.state_dict().items()] == keys['unet2d']
    got3 = [n for n, _ in n3.state_dict().items()]
    assert got3 == [n for n, _ in keys['unet3d']]
    with contextlib.redirect_stdout(io.StringIO()):
        makehebbian(n2, exclude=workloads.EXCLUDE_2D, hebb_params={'mode': 'swta_t', 'k': 50., 'alpha': 1.})
        makehebbian(n3, exclude=workloads.EXCLUDE_3D, hebb_params={'mode': 'swta_t', 'k': 50., 'alpha': 1.})
    h2 = [m for m in n2.modules() if hasattr(m, 'local_update')]
    h3 = [m for m in n3.modules() if hasattr(m, 'local_update')]
    assert len(h2) == 22 and all(type(m).__name__ == 'HebbianConv2d' and m.mode == 'swta' for m in h2)
    assert len(h3) == 22 and sum(type(m).__name__ == 'HebbianConvTranspose3d' for m in h3) == 4
    assert all(p.requires_grad for p in n2.out_conv.parameters())
    assert not any(p.requires_grad for n, p in n2.named_parameters() if 'out_conv' not in n and p.dim() == 1)


def test_flatten_delta_w_keeps_shapes_and_aliases():
    from hebb.step import flatten_delta_w, hebbian_layers
    with contextlib.redirect_stdout(io.StringIO()):
        net = makehebbian(Net(), exclude=['clf.5'], hebb_params={'mode': 'swta_t', 'k': 5., 'alpha': 1.})
    before = {n: (tuple(b.shape), b.stride()) for n, b in net.named_buffers() if n.endswith('delta_w')}
    keys = list(net.state_dict().keys())
    flat = flatten_delta_w(net)
    after = {n: (tuple(b.shape), b.stride()) for n, b in net.named_buffers() if n.endswith('delta_w')}
    assert before == after and keys == list(net.state_dict().keys())
    flat.fill_(2.0)
    assert all(float(m.delta_w.min()) == 2.0 for m in hebbian_layers(net))
    net.up[0].delta_w.zero_()
    assert float(flat.sum()) < 2.0 * flat.numel()


def test_reference_checkpoint_interchange():
    """A checkpoint written by the REFERENCE modules in its save_snapshot() layout (utils.py:29-55) loads into the
    drop-in network built the way the reference's consumers do (test_2d.py:105-109: makehebbian, then
    load_state_dict), and the drop-in's own state_dict loads back bit-identically."""
    ck = torch.load(os.path.join(ROOT, 'tests', 'golden', 'ref_checkpoint.pth'), map_location='cpu', weights_only=True)
    assert set(ck) == {'model', 'threshold', 'hebb_params', 'excluded_layers'}
    hp = dict(ck['hebb_params'])
    with contextlib.redirect_stdout(io.StringIO()):
        net = makehebbian(Net(), exclude=ck['excluded_layers'], hebb_params=hp)
    missing, unexpected = net.load_state_dict(ck['model'], strict=True)
    assert not missing and not unexpected
    for k, v in ck['model'].items():
        assert torch.equal(net.state_dict()[k], v), k
    up = net.up[0]
    assert up.weight.shape == (16, 20, 3, 3) and up.weight.stride() == (9, 144, 3, 1)       # view kept
    assert torch.equal(up.delta_w, ck['model']['up.0.delta_w'])
    with contextlib.redirect_stdout(io.StringIO()):
        twin = makehebbian(Net(), exclude=ck['excluded_layers'], hebb_params=hp)
    twin.load_state_dict(net.state_dict())
    assert all(torch.equal(a, b) for a, b in zip(twin.state_dict().values(), net.state_dict().values()))


def test_fused_head_bias_is_added_exactly_once_whoever_takes_it():
    """hebb.fused: the convolution alone decides whether it leaves its bias to the fused activation that follows and
    records that per call; the follower acts on the record (round-1 advisor finding: the two used to re-derive the
    decision from different tensors, and a disagreement dropped the bias silently)."""
    import copy
    from hebb.fused import fuse_norm_act, FusedBiasReluDropout, NoBiasFastWgradConv2d
    torch.manual_seed(0)
    net = nn.Module()
    net.head = nn.Sequential(nn.Conv2d(4, 8, 3, padding=1), nn.ReLU(), nn.Dropout(0.0), nn.Conv2d(8, 2, 1))
    ref = copy.deepcopy(net.head)
    keys = list(net.state_dict().keys())
    fuse_norm_act(net)
    assert list(net.state_dict().keys()) == keys
    conv, act = net.head[0], net.head[1]
    assert isinstance(conv, NoBiasFastWgradConv2d) and isinstance(act, FusedBiasReluDropout)
    x = torch.randn(2, 4, 6, 6)
    net.train(); ref.train()
    # CPU input: the convolution keeps its bias, the follower takes the stock ops and owes nothing
    assert torch.allclose(net.head(x), ref(x), atol=1e-6) and act._bias_pending is False
    # the convolution leaves its bias out (what it does for CUDA fp32 training inputs) while the follower's own input
    # cannot take the fused kernel (here: a CPU tensor): the follower adds the bias itself, once
    act._wants_bias = lambda t: True
    assert torch.allclose(net.head(x), ref(x), atol=1e-6) and act._bias_pending is False
    # and a follower that is called without its convolution having run first adds nothing
    z = torch.randn(2, 8, 6, 6)
    assert torch.allclose(act(z), torch.relu(z))


def test_flat_gradient_buffer_keeps_each_parameters_dimension_order():
    """HebbianStepper: the gradients of the back-prop parameters alias ONE flat buffer and keep their parameter's
    strides (channels_last head weights, transposed-view weights) -- fused optimisers insist on matching layouts."""
    from hebb.step import flatten_grads
    a = nn.Parameter(torch.randn(8, 4, 3, 3).contiguous(memory_format=torch.channels_last))
    b = nn.Parameter(torch.randn(6, 5, 2, 2).transpose(0, 1))
    c = nn.Parameter(torch.randn(7))
    flat = flatten_grads([a, b, c])
    for p in (a, b, c):
        assert p.grad.shape == p.shape and p.grad.stride() == p.stride()
        assert p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
    (a.sum() * 2 + (b * b.detach()).sum() + c.sum() * 3).backward()
    assert torch.equal(a.grad, torch.full_like(a, 2.0)) and torch.equal(c.grad, torch.full_like(c, 3.0))
    assert torch.allclose(b.grad, b.detach())
    assert float(flat.abs().sum()) > 0
