"""Pin the CPU oracle against outputs of the live reference (tests/golden/*)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import hebb_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
META = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'makehebbian_golden.json')))['meta']
CONV = [n for n, m in META.items() if m['kind'] == 'conv' and n != 'zero_row']
CONVT = [n for n, m in META.items() if m['kind'] == 'convT']
HPCA = [n for n, m in META.items() if m['kind'] == 'hpca']


def relerr(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _tup(v, nd):
    return tuple(v) if isinstance(v, (list, tuple)) else (v,) * nd


@pytest.mark.parametrize('name', CONV)
def test_conv_forward_and_delta(golden, name):
    m = META[name]
    nd = m['nd']
    x = torch.from_numpy(golden[name + '/x'])
    w = torch.from_numpy(golden[name + '/w'])
    b = torch.from_numpy(golden[name + '/b']) if (name + '/b') in golden else torch.zeros(m['Cout'])
    pad = m['padding'] if isinstance(m['padding'], int) else tuple(m['padding'])
    stride = _tup(m['stride'], nd)
    xp = O.zero_halo(x, pad, nd)
    y = O.conv_activation(xp, w, b, stride)
    assert relerr(y, golden[name + '/y']) < 2e-6
    assert np.array_equal(O.winners(y).numpy().astype(np.int32), golden[name + '/win'])
    dw = O.swta_delta(xp, y, w, m['k'], stride)
    # oracle and reference differ only in fp32 summation order
    assert relerr(dw, golden[name + '/dw1']) < 5e-6
    if (name + '/dw2') in golden:
        xp2 = O.zero_halo(x * 0.5, pad, nd)
        y2 = O.conv_activation(xp2, w, b, stride)
        dw2 = dw + O.swta_delta(xp2, y2, w, m['k'], stride)
        assert relerr(dw2, golden[name + '/dw2']) < 5e-6
        grad, z = O.fold_delta_into_grad(None, dw2, 1.0)
        assert relerr(grad, golden[name + '/grad']) < 5e-6
        assert float(z.abs().max()) == 0.0


@pytest.mark.parametrize('name', CONVT)
def test_convT_forward_and_delta(golden, name):
    m = META[name]
    nd = m['nd']
    x = torch.from_numpy(golden[name + '/x'])
    w = torch.from_numpy(golden[name + '/w'])
    stride = _tup(m['stride'], nd)
    y = O.convT_activation(x, w, None, stride)
    assert relerr(y, golden[name + '/y']) < 2e-6
    assert np.array_equal(O.winners(y).numpy().astype(np.int32), golden[name + '/win'])
    dw = O.swta_t_delta(x, y, w, m['k'], stride)
    assert relerr(dw, golden[name + '/dw1']) < 5e-6
    grad, _ = O.fold_delta_into_grad(None, dw, 1.0)
    assert relerr(grad, golden[name + '/grad']) < 5e-6


@pytest.mark.parametrize('name', HPCA)
def test_hpca_delta(golden, name):
    m = META[name]
    nd = m['nd']
    x, w, b = (torch.from_numpy(golden[name + s]) for s in ('/x', '/w', '/b'))
    xp = O.zero_halo(x, m['padding'], nd)
    y = O.conv_activation(xp, w, b, (1,) * nd)
    assert relerr(y, golden[name + '/y']) < 2e-6
    assert relerr(O.hpca_delta(xp, y, w, (1,) * nd), golden[name + '/dw1']) < 5e-6


@pytest.mark.parametrize('name', [n for n, m in META.items() if m['kind'] == 'hpcaT'])
def test_hpca_on_transposed_layer(golden, name):
    m = META[name]
    x, w = torch.from_numpy(golden[name + '/x']), torch.from_numpy(golden[name + '/w'])
    st = (2,) * m['nd']
    y = O.convT_activation(x, w, None, st)
    assert relerr(y, golden[name + '/y']) < 2e-6
    assert relerr(O.hpca_exchanged_delta(x, y, w, st), golden[name + '/dw1']) < 5e-6


def test_fp64_oracle_agrees_with_fp32_golden(golden):
    """The dtype-agnostic oracle in fp64 bounds the reference's own fp32 noise."""
    name = 'c2d_c1_config1'
    m = META[name]
    x = torch.from_numpy(golden[name + '/x']).double()
    w = torch.from_numpy(golden[name + '/w']).double()
    xp = O.zero_halo(x, m['padding'], 2)
    y = O.conv_activation(xp, w, None, (1, 1))
    dw = O.swta_delta(xp, y, w, m['k'], (1, 1))
    assert relerr(dw, golden[name + '/dw1']) < 5e-6
    assert np.array_equal(O.winners(y).numpy().astype(np.int32), golden[name + '/win'])


@pytest.mark.parametrize('opt_name', ['sgd', 'adam'])
def test_hundred_step_drift(golden, opt_name):
    """W after 1 and 100 optimiser steps, driven exactly like the reference loop
    (pretrain_hebbian_unsup_2d.py:181-196 without the back-prop head)."""
    xs = torch.from_numpy(golden[f'drift_{opt_name}/xs'])
    layer = O.OracleHebbConv(2, 3, 16, 3, padding=1, bias=False, k=10., alpha=1.)
    with torch.no_grad():
        layer.weight.copy_(torch.from_numpy(golden[f'drift_{opt_name}/w0']))
    opt = (torch.optim.SGD([layer.weight], lr=1e-3) if opt_name == 'sgd'
           else torch.optim.Adam([layer.weight], lr=1e-3))
    layer.train()
    for step in range(100):
        opt.zero_grad()
        layer(xs[step % 4])
        layer.local_update()
        opt.step()
        if step == 0:
            assert relerr(layer.weight.detach(), golden[f'drift_{opt_name}/w1']) < 1e-6
    assert relerr(layer.weight.detach(), golden[f'drift_{opt_name}/w100']) < 1e-4


def test_zero_norm_row_and_no_update_modes(golden):
    x = torch.from_numpy(golden['zero_row/x'])
    w = torch.from_numpy(golden['zero_row/w'])
    xp = O.zero_halo(x, 1, 2)
    y = O.conv_activation(xp, w, None, (1, 1))
    assert torch.isfinite(y).all()
    assert relerr(y, golden['zero_row/y']) < 2e-6
    assert float(y[:, 2].abs().max()) == 0.0
    assert relerr(O.swta_delta(xp, y, w, 5., (1, 1)), golden['zero_row/dw1']) < 5e-6
    layer = O.OracleHebbConv(2, 3, 4, 3, padding=1, bias=False, k=5., alpha=1.)
    layer.eval()
    layer(x)
    assert float(layer.delta_w.abs().max()) == 0.0
    layer.train()
    layer.alpha = 0.
    layer(x)
    assert float(layer.delta_w.abs().max()) == 0.0


def test_identities():
    """Known-answer identities: sum_c r = 1, batch-shard additivity,
    and the Hebbian term == conv weight-gradient with r in place of dL/dy."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 5, 9, 8, generator=g, dtype=torch.float64)
    w = torch.randn(7, 5, 3, 3, generator=g, dtype=torch.float64)
    xp = O.zero_halo(x, 1, 2)
    y = O.conv_activation(xp, w, None, (1, 1))
    r = O.swta_response(y, 4.0)
    assert torch.allclose(r.sum(0), torch.ones(r.shape[1], dtype=torch.float64))
    full = O.swta_delta(xp, y, w, 4.0, (1, 1))
    parts = sum(O.swta_delta(xp[i:i + 2], y[i:i + 2], w, 4.0, (1, 1)) for i in (0, 2))
    assert relerr(parts, full) < 1e-12
    rr = (y * 4.0).softmax(dim=1)
    hebb = torch.nn.grad.conv2d_weight(xp, w.shape, rr)
    dec = r.sum(1).reshape(-1, 1, 1, 1) * w
    assert relerr(hebb - dec, full) < 1e-12


@pytest.mark.parametrize('name', [n for n, m in META.items() if m['kind'] == 'contrastive'])
def test_contrastive_delta(golden, name):
    """hebb.py:143-172 — the permutation the reference drew is part of the fixture."""
    m = META[name]
    x, w, b = (torch.from_numpy(golden[name + k]) for k in ('/x', '/w', '/b'))
    xp = O.zero_halo(x, m['padding'], m['nd'])
    y = O.conv_activation(xp, w, b, (1,) * m['nd'])
    assert relerr(y, golden[name + '/y']) < 2e-6
    gw, gb = O.contrastive_delta(xp, w, b, (1,) * m['nd'], m['contrast'], perm=torch.from_numpy(golden[name + '/perm']))
    assert relerr(gw, golden[name + '/dw1']) < 5e-6
    if m['bias']:
        assert relerr(gb, golden[name + '/gb']) < 5e-6
