"""Whole-network pinning of the oracle (module wrappers + restated topologies) against the live
reference (tests/golden/network_golden.json), and the data-parallel exchange on CPU/gloo."""
import io
import contextlib
import json
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import workloads
from oracle import hebb_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NET = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'network_golden.json')))


from helpers import digest_err  # noqa: E402


def build_oracle_net(name):
    if name == 'unet2d':
        net, excl = workloads.unet2d(3, 2), workloads.EXCLUDE_2D
    else:
        net, excl = workloads.UNet3D(1, 2, init_features=4), workloads.EXCLUDE_3D
    O.oracle_makehebbian(net, exclude=excl, k=50., alpha=1.)
    workloads.deterministic_state_(net)
    workloads.disable_dropout_(net)
    return net.train()


@pytest.mark.parametrize('name', ['unet2d', 'unet3d_f4'])
def test_oracle_network_matches_reference(name):
    net = build_oracle_net(name)
    gold = NET[name]
    x = torch.randn(*gold['shape'], generator=torch.Generator().manual_seed(77))
    out = net(x)
    assert digest_err(out, gold['out']) < 1e-4
    layers = {n: m for n, m in net.named_modules() if hasattr(m, 'local_update')}
    assert sorted(layers) == sorted(gold['layers'])
    for n, m in layers.items():
        assert digest_err(m.delta_w, gold['layers'][n]['delta_w']) < 2e-4, n
        m.local_update()
        assert digest_err(m.weight.grad, gold['layers'][n]['grad']) < 2e-4, n


def _rank_main(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.join(ROOT, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200'))
        from hebb.step import flatten_delta_w, HebbianStepper, hebbian_layers
        torch.manual_seed(0)
        net = torch.nn.Sequential(O.OracleHebbConv(2, 3, 8, 3, padding=1, bias=False, k=5., alpha=1.),
                                  O.OracleHebbConvT(2, 8, 4, 2, stride=2, bias=False, k=5., alpha=1.))
        net.train()
        x = torch.randn(4, 3, 8, 8, generator=torch.Generator().manual_seed(5))
        ref = [None, None]
        if rank == 0:                       # full-batch answer
            net(x)
            ref = [m.delta_w.clone() for m in net]
            for m in net:
                m.delta_w.zero_()
        opt = torch.optim.SGD(net.parameters(), lr=0.0)
        st = HebbianStepper(net, opt)
        assert st.allreduce and st.flat is not None
        # the second layer's input depends only on its own shard, so per-layer additivity holds
        net(x[rank * 2:(rank + 1) * 2])
        st.exchange()                        # ONE all-reduce for both layers
        if rank == 0:
            errs = [float((m.delta_w - r).norm() / r.norm()) for m, r in zip(net, ref)]
            q.put(errs)
    finally:
        dist.destroy_process_group()


def test_batch_shard_allreduce_gloo_world2():
    """ΔW(batch) == all-reduce-sum of ΔW(shards): one collective over the flat buffer."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    errs = q.get(timeout=5)
    assert max(errs) < 1e-5, errs
