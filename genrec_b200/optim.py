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

from typing import Optional

import torch
import torch.distributed as dist

from . import functional as Fn


class FlatBuffers:
    """Device-agnostic part: one flat fp32 parameter buffer, one flat gradient buffer, one flat bf16 mirror, with every
    ``param.data`` / ``param.grad`` re-homed as a view (256-byte aligned slots).  Works on CPU tensors too, which is how the
    world_size-2 gloo tests exercise the data-parallel plumbing without a GPU."""

    def __init__(self, model: torch.nn.Module):
        params = [p for p in model.parameters() if p.requires_grad]
        assert params, "no trainable parameters"
        dev = params[0].device
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 63) // 64 * 64          # 256-byte aligned slots (bf16 views stay 16-byte aligned)
        self.n, self.offsets, self.params, self.device = n, offs, params, dev
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.mirror = torch.zeros(n, dtype=torch.bfloat16, device=dev)
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


class FlatAdam:
    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, process_group=None, grad_sink: bool = True, unit_loss_grad: bool = False):
        """``unit_loss_grad=True`` promises that every training forward is followed by exactly one ``loss.backward()`` with
        gradient 1 (no loss scaling, no gradient accumulation through a scaled loss): the fused head then accumulates its
        parameter gradients into the flat buffer in the same pass that computes the loss.  The promise is checked on the
        device (``grb_assert_unit_scalar``).  Default False: fully general, a few microseconds slower per step."""
        self.buffers = FlatBuffers(model)
        dev = self.buffers.device
        assert dev.type == "cuda", "FlatAdam drives CUDA kernels; move the model to the GPU first"
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

    def mirror_of(self, p: torch.Tensor) -> torch.Tensor:
        return self.buffers.mirror_of(p)

    def grad_of(self, p: torch.Tensor) -> Optional[torch.Tensor]:
        return self.buffers.grad_of(p)

    def world(self) -> int:
        return world_size(self.group)

    def step(self) -> None:
        """all-reduce(SUM) -> fused Adam with grad_scale = 1/world -> mirror refresh -> grad zeroed."""
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
        return {"m": self.m.clone(), "v": self.v.clone(), "state": self.state.clone(),
                "hyper": dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.weight_decay), "n": self.n}

    def load_state_dict(self, sd: dict) -> None:
        if int(sd["n"]) != self.n:
            raise ValueError(f"FlatAdam state for {sd['n']} elements does not fit this model ({self.n})")
        self.m.copy_(sd["m"]); self.v.copy_(sd["v"]); self.state.copy_(sd["state"])
        h = sd.get("hyper", {})
        self.lr, self.betas = h.get("lr", self.lr), tuple(h.get("betas", self.betas))
        self.eps, self.weight_decay = h.get("eps", self.eps), h.get("weight_decay", self.weight_decay)
        self.refresh_mirror()
