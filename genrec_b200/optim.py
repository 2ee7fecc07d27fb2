"""Flat-buffer training runtime: fused Adam + bf16 weight mirror + one NCCL all-reduce per step.

``FlatAdam(model)`` re-homes every parameter of the model as a view into ONE contiguous fp32 buffer, gives each a
``.grad`` view into one flat gradient buffer, and keeps a flat bf16 mirror that the CUDA kernels read as tensor-core
operands.  Consequences ("distributed communication backend"):
  * backward kernels accumulate weight gradients straight into the flat buffer (``grad sink``) - no per-parameter
    AccumulateGrad kernels, no flatten copy before the all-reduce;
  * data-parallel training = ONE ``all_reduce(SUM)`` over the flat gradient (NCCL over NVLink/NVSwitch), the 1/world
    scale folded into the fused Adam kernel, which also refreshes the bf16 mirror and zeroes the gradient;
  * everything is CUDA-graph capturable (Adam's bias correction is ticked on the device).
Semantics = ``torch.optim.Adam(params, lr, betas, eps, weight_decay)`` followed by ``zero_grad()``.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from . import functional as Fn
from ._lib import check, ptr, stream_ptr


class FlatBuffers:
    """Device-agnostic part: one flat fp32 parameter buffer, one flat gradient buffer, one flat bf16 mirror, with every
    ``param.data`` / ``param.grad`` re-homed as a view (256-byte aligned slots).  Works on CPU tensors too, which is how the
    world_size-2 gloo tests exercise the data-parallel plumbing without a GPU."""

    def __init__(self, model: torch.nn.Module, alloc=None, pad_to: int = 64):
        """alloc(n, dtype) -> zeroed 1-D tensor (default torch.zeros on the parameters' device); pad_to: the total length is
        rounded up to a multiple of it (the peer-memory optimizer needs 8 * world)."""
        params = [p for p in model.parameters() if p.requires_grad]
        assert params, "no trainable parameters"
        dev = params[0].device
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 63) // 64 * 64          # 256-byte aligned slots (bf16 views stay 16-byte aligned)
        n = (n + pad_to - 1) // pad_to * pad_to
        self.n, self.offsets, self.params, self.device = n, offs, params, dev
        if alloc is None:
            alloc = lambda k, dt: torch.zeros(k, dtype=dt, device=dev)  # noqa: E731
        self.flat = alloc(n, torch.float32)
        self.grad = alloc(n, torch.float32)
        self.mirror = alloc(n, torch.bfloat16)
        self._mirror_view, self._grad_view = {}, {}
        with torch.no_grad():
            for p, o in zip(params, offs):
                k = p.numel()
                self.flat[o:o + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[o:o + k].view(p.shape)
                p.grad = self.grad[o:o + k].view(p.shape)
                self._mirror_view[id(p)] = self.mirror[o:o + k].view(p.shape)
                self._grad_view[id(p)] = p.grad

    def mirror_of(self, p: torch.Tensor) -> torch.Tensor:
        return self._mirror_view[id(p)]

    def grad_of(self, p: torch.Tensor) -> Optional[torch.Tensor]:
        return self._grad_view.get(id(p))


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_gradients(buffers: FlatBuffers, group=None) -> float:
    """ONE all-reduce(SUM) over the flat gradient buffer; returns the scale (1/world) the optimizer must apply, i.e. DDP's
    gradient averaging (each rank back-propagates its own mean-over-valid-tokens loss)."""
    w = world_size(group)
    if w > 1:
        dist.all_reduce(buffers.grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / w


class PeerMemory:
    """Symmetric (peer-mapped, multicast-mapped where the fabric supports it) allocations for the one-pass data-parallel optimizer
    step ``grb_dp_adam_step`` (csrc/dp_adam.cuh).  Built on ``torch.distributed._symmetric_memory`` for the rendezvous only -
    the collective itself is our kernel.  Raises if symmetric memory is unavailable; FlatAdam then keeps the NCCL all-reduce."""

    def __init__(self, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.symm, self.device = symm, device
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.handles = {}
        self.tensors = []

    def alloc(self, n: int, dtype) -> torch.Tensor:
        t = self.symm.empty(n, dtype=dtype, device=self.device)
        h = self.symm.rendezvous(t, self.group.group_name)
        t.zero_()
        self.handles[t.data_ptr()] = h
        self.tensors.append(t)
        return t

    def handle(self, t: torch.Tensor):
        return self.handles[t.data_ptr()]

    def peer_ptrs(self, t: torch.Tensor) -> torch.Tensor:
        return torch.tensor([int(x) for x in self.handle(t).buffer_ptrs], dtype=torch.int64, device=self.device)

    def multicast_ptr(self, t: torch.Tensor) -> int:
        h = self.handle(t)
        try:
            return int(h.multicast_ptr) if h.has_multicast_support(self.device.type, self.device.index) or int(h.multicast_ptr) else 0
        except Exception:  # noqa: BLE001
            return int(getattr(h, "multicast_ptr", 0) or 0)


class FlatAdam:
    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, process_group=None, grad_sink: bool = True, unit_loss_grad: bool = False,
                 peer_memory: Optional[bool] = None, defer_weight_grads: bool = False):
        """``unit_loss_grad=True`` promises that every training forward is followed by exactly one ``loss.backward()`` with
        gradient 1 (no loss scaling, no gradient accumulation through a scaled loss): the fused head then accumulates its
        parameter gradients into the flat buffer in the same pass that computes the loss.  The promise is checked on the
        device (``grb_assert_unit_scalar``).  Default False: fully general, a few microseconds slower per step.
        ``defer_weight_grads=True`` moves the dW / dE GEMMs of the backward pass to a side stream that ``step()`` joins (they are not
        on the critical path); anything else that reads ``.grad`` / the flat gradient before ``step()`` must call
        ``sync_grads()`` first.  Process-wide switch (``grb_set_defer_weight_grads``)."""
        dev0 = next(p for p in model.parameters() if p.requires_grad).device
        assert dev0.type == "cuda", "FlatAdam drives CUDA kernels; move the model to the GPU first"
        # data-parallel step: one pass over NVLink peer memory (multimem reduce + Adam + multicast parameter store, csrc/dp_adam.cuh)
        # when symmetric memory can be set up, otherwise NCCL all-reduce + the local fused Adam.  GRB_DP=nccl forces the latter.
        self.peer = None
        self.dp_mode = "single"
        w = world_size(process_group)
        want_peer = (os.environ.get("GRB_DP", "") != "nccl") if peer_memory is None else bool(peer_memory)
        if w > 1:
            self.dp_mode = "nccl-allreduce"
            if want_peer:
                try:
                    self.peer = PeerMemory(dev0, process_group)
                    self.buffers = FlatBuffers(model, alloc=self.peer.alloc, pad_to=64 * w)
                    self._peer_setup()
                    self.dp_mode = "peer-multimem" if self._mc[0] else "peer-p2p"
                except Exception as e:  # noqa: BLE001
                    if peer_memory:
                        raise
                    self.peer = None
                    self._peer_error = f"{type(e).__name__}: {e}"
        if self.peer is None:
            self.buffers = FlatBuffers(model)
        dev = self.buffers.device
        self.model, self.lr, self.betas, self.eps, self.weight_decay = model, lr, betas, eps, weight_decay
        self.group = process_group
        self.n, self.flat, self.grad, self.mirror, self.params = (self.buffers.n, self.buffers.flat, self.buffers.grad,
                                                                    self.buffers.mirror, self.buffers.params)
        self.m = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.state = torch.zeros(4, dtype=torch.float32, device=dev)
        Fn.cast_bf16(self.flat, self.mirror)
        # hand the mirror (and, optionally, the gradient sink) to the modules
        for mod in model.modules():
            if hasattr(mod, "_bf16_provider"):
                mod._bf16_provider = self.buffers.mirror_of
            if grad_sink and hasattr(mod, "_grad_sink"):
                mod._grad_sink = self.buffers.grad_of
            if grad_sink and hasattr(mod, "_unit_loss_grad"):
                mod._unit_loss_grad = bool(unit_loss_grad)
        # the kernels read the bf16 mirror, never the fp32 masters: anything that rewrites the masters behind the optimizer's
        # back (load_state_dict on resume, accelerate.load_state) must refresh it
        self._hook = model.register_load_state_dict_post_hook(lambda module, incompatible: self.refresh_mirror())
        if defer_weight_grads:
            assert grad_sink, "deferred weight gradients need the gradient sink (nothing but this optimizer consumes them)"
            Fn.set_defer_weight_grads(True)

    def mirror_of(self, p: torch.Tensor) -> torch.Tensor:
        return self.buffers.mirror_of(p)

    def grad_of(self, p: torch.Tensor) -> Optional[torch.Tensor]:
        return self.buffers.grad_of(p)

    def world(self) -> int:
        return world_size(self.group)

    def _peer_setup(self) -> None:
        pm, b = self.peer, self.buffers
        self._sig = pm.alloc(2 * pm.world, torch.int32)
        self._epoch = torch.zeros(2, dtype=torch.int32, device=b.device)
        self._peer_ptrs = tuple(pm.peer_ptrs(t) for t in (b.grad, b.flat, b.mirror, self._sig))
        self._mc = tuple(pm.multicast_ptr(t) for t in (b.grad, b.flat, b.mirror))
        if not all(self._mc):
            self._mc = (0, 0, 0)
        torch.cuda.synchronize(b.device)
        dist.barrier(pm.group)

    def sync_grads(self) -> None:
        """Wait (on the current stream) for gradient work that was deferred to the side stream."""
        Fn.join_deferred(self.flat.device)

    def step(self) -> None:
        """world == 1: fused Adam.  world > 1: reduce + Adam + parameter broadcast in one pass over peer memory (dp_mode "peer-*"),
        or all-reduce(SUM) -> fused Adam with grad_scale = 1/world (dp_mode "nccl-allreduce"); either way the bf16 mirror is
        refreshed and the flat gradient is zero afterwards."""
        Fn.join_deferred(self.flat.device)
        if self.peer is not None:
            pg, pp, pmir, psig = self._peer_ptrs
            with torch.cuda.device(self.flat.device):
                check(_lib.load().grb_dp_adam_step(
                    ptr(self.flat), ptr(self.grad), ptr(self.m), ptr(self.v), ptr(self.mirror), self._mc[0] or None, self._mc[1] or None,
                    self._mc[2] or None, ptr(pg), ptr(pp), ptr(pmir), ptr(psig), ptr(self._sig), ptr(self._epoch), self.n, self.peer.rank,
                    self.peer.world, ptr(self.state), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                    1.0 / self.peer.world, stream_ptr(self.flat.device)))
            return
        scale = allreduce_gradients(self.buffers, self.group)
        Fn.adam_step(self.flat, self.grad, self.m, self.v, self.mirror, self.state, self.lr, self.betas[0], self.betas[1],
                     self.eps, self.weight_decay, scale, True)

    def zero_grad(self, set_to_none: bool = False) -> None:
        pass  # the fused step already zeroed the flat gradient

    def refresh_mirror(self) -> None:
        """Re-derive the bf16 operand mirror from the fp32 masters.  Called automatically after ``model.load_state_dict``;
        call it yourself after editing ``param.data`` in place (manual re-initialisation)."""
        Fn.cast_bf16(self.flat, self.mirror)

    def state_dict(self) -> dict:
        """Adam moments + step state (``torch.optim.Adam``-style checkpointing; parameters live in ``model.state_dict()``)."""
        m, v = self.m.clone(), self.v.clone()
        if self.peer is not None:     # each rank holds the moments of its own slice only: assemble the full vectors
            per = self.n // self.peer.world
            lo = self.peer.rank * per
            dist.all_gather_into_tensor(m, self.m[lo:lo + per].clone(), group=self.peer.group)
            dist.all_gather_into_tensor(v, self.v[lo:lo + per].clone(), group=self.peer.group)
        return {"m": m, "v": v, "state": self.state.clone(),
                "hyper": dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.weight_decay), "n": self.n}

    def load_state_dict(self, sd: dict) -> None:
        if int(sd["n"]) != self.n:
            raise ValueError(f"FlatAdam state for {sd['n']} elements does not fit this model ({self.n})")
        self.m.copy_(sd["m"]); self.v.copy_(sd["v"]); self.state.copy_(sd["state"])
        h = sd.get("hyper", {})
        self.lr, self.betas = h.get("lr", self.lr), tuple(h.get("betas", self.betas))
        self.eps, self.weight_decay = h.get("eps", self.eps), h.get("weight_decay", self.weight_decay)
        self.refresh_mirror()
