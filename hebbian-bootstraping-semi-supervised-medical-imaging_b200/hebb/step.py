"""Step-level helpers around the Hebbian layers (not part of the reference API).

The reference's training loop (pretrain_hebbian_unsup_2d.py:181-196) does, per batch,
    optimizer.zero_grad(); out = model(x); loss.backward()
    for m in model.modules():  m.local_update()      # if it has one
    optimizer.step()
`HebbianStepper` restates that step with B200-side changes that keep results identical:
  * all delta_w buffers alias ONE flat fp32 buffer, so the data-parallel exchange of the plasticity
    updates is a single all-reduce(sum) per step (the update is a sum over patches, hence additive over
    batch shards — SURVEY.md §3.5-vi); it is issued asynchronously right after the forward pass, so it
    overlaps the backward pass of the back-prop head,
  * the gradients of every back-prop parameter (the `exclude`d head of makehebbian, layers with
    alpha < 1, trainable biases) alias a second flat buffer that is averaged over the ranks after the
    backward pass (what DistributedDataParallel would do), so all replicas take the same optimiser step,
  * local_update() of every layer runs as one multi-tensor kernel launch,
  * optionally the whole step is replayed from a CUDA graph (`capture=True`).
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
        if getattr(m, '_transposed', False) or (not old.is_contiguous() and old.transpose(0, 1).is_contiguous()):
            base = flat[o:o + n].view(old.shape[1], old.shape[0], *old.shape[2:])
            base.copy_(old.transpose(0, 1))
            new = base.transpose(0, 1)
        else:
            new = flat[o:o + n].view(old.shape)
            new.copy_(old)
        m._buffers['delta_w'] = new
    return flat


def backprop_parameters(model: nn.Module, layers: Optional[List[nn.Module]] = None) -> List[nn.Parameter]:
    """Parameters whose .grad comes out of loss.backward(): everything trainable except the weights of fully
    Hebbian layers (alpha == 1: their forward detaches the weight, hebb/_core.py, so autograd never reaches it and
    weight.grad is produced by local_update() alone)."""
    layers = hebbian_layers(model) if layers is None else layers
    hebb_only = {id(m.weight) for m in layers if getattr(m, 'alpha', 0) == 1}
    out, seen = [], set()
    for p in model.parameters():
        if p.requires_grad and id(p) not in hebb_only and id(p) not in seen:
            seen.add(id(p))
            out.append(p)
    return out


def _dense_permutation(t: torch.Tensor) -> bool:
    """True if t's elements tile its storage extent exactly once (a permuted contiguous tensor)."""
    dims = sorted((st, sz) for st, sz in zip(t.stride(), t.shape) if sz > 1)
    run = 1
    for st, sz in dims:
        if st != run:
            return False
        run *= sz
    return True


def flatten_grads(params: List[nn.Parameter], align: int = 64) -> Optional[torch.Tensor]:
    """Point the .grad of every parameter into one flat zero-initialised buffer (dense parameters keep their
    strides, whatever the dimension order; anything else gets a contiguous gradient).  autograd accumulates in place into an existing .grad,
    so the buffer IS the gradient after backward()."""
    if not params:
        return None
    offs, total = [], 0
    for p in params:
        offs.append(total)
        total += (p.numel() + align - 1) // align * align
    flat = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
    for p, o in zip(params, offs):
        n = p.numel()
        if p.is_contiguous() or not _dense_permutation(p):
            g = flat[o:o + n].view(p.shape)
        else:
            # same strides as the parameter (channels_last head weights, the transposed-view weights of
            # HebbianConvTranspose): what autograd would allocate, and what fused optimisers insist on
            g = flat[o:o + n].as_strided(p.shape, p.stride())
        p.grad = g
    return flat


@torch.no_grad()
def local_update_all(layers: List[nn.Module]):
    """Every layer's local_update() (hebb.py:174-192) in one kernel launch."""
    grads, dws, alphas, had = [], [], [], []
    for m in layers:
        dw = m.delta_w
        if not dw.is_cuda:            # a layer that is not ours (the CPU oracle in the tests): its own method
            m.local_update()
            continue
        h = m.weight.grad is not None
        if not h:
            m.weight.grad = torch.empty_like(dw)
        g = m.weight.grad
        if g.stride() != dw.stride():
            m.local_update()          # odd layout: per-layer path handles it
            continue
        grads.append(g); dws.append(dw); alphas.append(m.alpha); had.append(h)
    if dws:
        _native.local_update_multi(grads, dws, alphas, had)


class HebbianStepper:
    """One Hebbian pretraining step (the unit samples/s counts), optionally data-parallel.

    process_group / allreduce: the data-parallel exchange (default: on iff torch.distributed is initialised with
    more than one rank).  overlap: issue the delta_w all-reduce asynchronously after the forward pass.
    capture: record the step (zero gradients, forward, both collectives, loss, backward, local_update, optimiser
    step) into a CUDA graph on first use and replay it afterwards; needs static input buffers (step() copies into
    them) and a capturable optimiser.  `graph_launches` counts the library's kernel launches replayed so far (the
    library's own counter only sees the launches made while recording)."""

    def __init__(self, model: nn.Module, optimizer: torch.optim.Optimizer, criterion=None,
                 process_group=None, allreduce: Optional[bool] = None, overlap: bool = True, capture: bool = False):
        self.model = model
        self.optimizer = optimizer
        self.criterion = criterion
        self.group = process_group
        self.layers = hebbian_layers(model)
        self.flat = flatten_delta_w(model)
        if allreduce is None:
            allreduce = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        self.allreduce = allreduce
        self.overlap = overlap
        self.world = dist.get_world_size(process_group) if (allreduce and dist.is_initialized()) else 1
        self.bp_params = backprop_parameters(model, self.layers)
        self.hebb_only = [m for m in self.layers if getattr(m, 'alpha', 0) == 1]
        # the back-prop gradients live in one flat buffer: one collective, and zeroing them is one fill
        self.flat_grad = flatten_grads(self.bp_params)
        self._grad_views = [p.grad for p in self.bp_params] if self.flat_grad is not None else []
        self._pending = None
        self.capture = capture
        self._graph = None
        self._static = None
        self.graph_launches = 0
        self._launches_per_replay = 0

    # ------------------------------------------------------------------ data-parallel exchange
    def exchange_begin(self):
        """Start summing the per-rank partial delta_w of ALL layers (one collective over the flat buffer)."""
        if self.allreduce and self.flat is not None:
            self._pending = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=self.overlap)

    def exchange_end(self):
        """Average the back-prop gradients over the ranks (one collective) and wait for the delta_w sum."""
        if self.allreduce and self.flat_grad is not None:
            self.flat_grad.mul_(1.0 / self.world)
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        if self._pending is not None:
            self._pending.wait()
        self._pending = None

    def exchange(self):
        """Both collectives back to back (kept for callers that drive the step themselves)."""
        self.exchange_begin()
        self.exchange_end()

    # ------------------------------------------------------------------ the step
    def _zero_grad(self):
        # == optimizer.zero_grad(): the back-prop gradients are zeroed in place (their storage is the flat buffer),
        # the fully Hebbian weights get their gradient from local_update() alone
        if self.flat_grad is not None:
            self.flat_grad.zero_()
            for p, g in zip(self.bp_params, self._grad_views):     # (a caller's own zero_grad() may have dropped the views)
                p.grad = g
        for m in self.hebb_only:
            m.weight.grad = None

    def _step_body(self, x, target):
        self._zero_grad()
        out = self.model(x)
        self.exchange_begin()                 # delta_w of every layer is final once the forward pass is enqueued
        loss = None
        if self.criterion is not None and target is not None:
            loss = self.criterion(out, target)
            if loss.requires_grad:
                loss.backward()
        self.exchange_end()
        local_update_all(self.layers)
        self.optimizer.step()
        return out, loss

    def step(self, x, target=None):
        if not self.model.training:          # forward-only use (e.g. throughput of the alpha=0 network)
            with torch.no_grad():
                return self.model(x), None
        if not self.capture:
            return self._step_body(x, target)
        return self._graph_step(x, target)

    # ------------------------------------------------------------------ CUDA-graph replay
    def _graph_step(self, x, target):
        if self._graph is None:
            self._static = (x.clone(), target.clone() if target is not None else None)
            sx, st = self._static
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):     # warm-up outside the capture (allocator, lazy initialisation)
                for _ in range(2):
                    self._step_body(sx, st)
            torch.cuda.current_stream(x.device).wait_stream(side)
            for m in self.hebb_only:          # a captured local_update must see the same gradient tensor every replay
                if m.weight.grad is None:
                    m.weight.grad = torch.zeros_like(m.delta_w)
            self._keep_hebb_grads = True
            g = torch.cuda.CUDAGraph()
            n0 = _native.launch_count()
            # thread_local: the NCCL watchdog thread polls events while we record
            with torch.cuda.graph(g, capture_error_mode='thread_local' if self.allreduce else 'global'):
                self._out = self._captured_body(sx, st)
            self._launches_per_replay = _native.launch_count() - n0
            self._graph = g
        sx, st = self._static
        if x.data_ptr() != sx.data_ptr():
            sx.copy_(x, non_blocking=True)
        if st is not None and target.data_ptr() != st.data_ptr():
            st.copy_(target, non_blocking=True)
        self._graph.replay()
        self.graph_launches += self._launches_per_replay
        return self._out

    def release(self):
        """Drop the recorded graph (the next step() records a new one).  Call it before
        torch.distributed.destroy_process_group(): NCCL keeps a communicator alive -- and its destruction waiting --
        for as long as a CUDA graph holds one of its collectives."""
        self._graph = None
        self._out = None
        self._static = None

    def static_inputs(self):
        """(x, target) buffers the captured step reads; a caller may fill them directly to save the copy in step()."""
        return self._static

    def _captured_body(self, x, target):
        # like _step_body, but the fully Hebbian weights keep one persistent gradient tensor that local_update
        # overwrites (grad = -alpha * delta_w): `had` must be False for them without dropping the tensor
        if self.flat_grad is not None:
            self.flat_grad.zero_()
        out = self.model(x)
        self.exchange_begin()
        loss = None
        if self.criterion is not None and target is not None:
            loss = self.criterion(out, target)
            if loss.requires_grad:
                loss.backward()
        self.exchange_end()
        grads, dws, alphas, had = [], [], [], []
        hebb_only = {id(m) for m in self.hebb_only}
        for m in self.layers:
            grads.append(m.weight.grad); dws.append(m.delta_w); alphas.append(m.alpha); had.append(id(m) not in hebb_only)
        _native.local_update_multi(grads, dws, alphas, had)
        self.optimizer.step()
        return out, loss
