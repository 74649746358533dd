"""Step-level helpers around the Hebbian layers (not part of the reference API).

The reference's training loop (pretrain_hebbian_unsup_2d.py:181-196) does, per batch,
    optimizer.zero_grad(); out = model(x); loss.backward()
    for m in model.modules():  m.local_update()      # if it has one
    optimizer.step()
`HebbianStepper` restates that step with two B200-side changes that keep results identical:
  * all delta_w buffers alias ONE flat fp32 buffer, so the data-parallel exchange is a single
    all-reduce(sum) per step (the update is a sum over patches, hence additive over batch
    shards — SURVEY.md §3.5-vi), and
  * local_update() of every layer runs as one multi-tensor kernel launch.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _native


def hebbian_layers(model: nn.Module) -> List[nn.Module]:
    """Layers discovered the way the reference does it: anything with a local_update()."""
    return [m for m in model.modules() if hasattr(m, 'local_update') and hasattr(m, 'delta_w')]


def flatten_delta_w(model: nn.Module, align: int = 64) -> Optional[torch.Tensor]:
    """Re-point every layer's `delta_w` buffer into one flat, zero-initialised fp32 tensor.

    Shapes, strides (incl. the transposed views) and state_dict keys are unchanged.  Call it after
    the model is on its final device.  Returns the flat buffer (None if there are no layers).
    """
    layers = hebbian_layers(model)
    if not layers:
        return None
    offs, total = [], 0
    for m in layers:
        offs.append(total)
        total += (m.delta_w.numel() + align - 1) // align * align
    dev = layers[0].delta_w.device
    flat = torch.zeros(total, dtype=torch.float32, device=dev)
    for m, o in zip(layers, offs):
        old = m.delta_w
        n = old.numel()
        if getattr(m, '_transposed', False):
            base = flat[o:o + n].view(old.shape[1], old.shape[0], *old.shape[2:])
            base.copy_(old.transpose(0, 1))
            new = base.transpose(0, 1)
        else:
            new = flat[o:o + n].view(old.shape)
            new.copy_(old)
        m._buffers['delta_w'] = new
    return flat


@torch.no_grad()
def local_update_all(layers: List[nn.Module]):
    """Every layer's local_update() (hebb.py:174-192) in one kernel launch."""
    grads, dws, alphas, had = [], [], [], []
    for m in layers:
        dw = m.delta_w
        h = m.weight.grad is not None
        if not h:
            m.weight.grad = torch.empty_like(dw)
        g = m.weight.grad
        if g.stride() != dw.stride():
            m.local_update()          # odd layout: per-layer path handles it
            continue
        grads.append(g); dws.append(dw); alphas.append(m.alpha); had.append(h)
    _native.local_update_multi(grads, dws, alphas, had)


class HebbianStepper:
    """One Hebbian pretraining step (the unit samples/s counts), optionally data-parallel."""

    def __init__(self, model: nn.Module, optimizer: torch.optim.Optimizer, criterion=None,
                 process_group=None, allreduce: Optional[bool] = None):
        self.model = model
        self.optimizer = optimizer
        self.criterion = criterion
        self.group = process_group
        self.layers = hebbian_layers(model)
        self.flat = flatten_delta_w(model)
        if allreduce is None:
            allreduce = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        self.allreduce = allreduce

    def exchange(self):
        """Sum the per-rank partial delta_w of ALL layers with one collective."""
        if self.allreduce and self.flat is not None:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)

    def step(self, x, target=None):
        if not self.model.training:          # forward-only use (e.g. throughput of the alpha=0 network)
            with torch.no_grad():
                return self.model(x), None
        self.optimizer.zero_grad()
        out = self.model(x)
        loss = None
        if self.criterion is not None and target is not None:
            loss = self.criterion(out, target)
            if loss.requires_grad:
                loss.backward()
        self.exchange()
        local_update_all(self.layers)
        self.optimizer.step()
        return out, loss
