"""Oracle restatement of the RQ-VAE residual nearest-codebook search
(reference: genrec/models/rqvae.py:176-199, :246-254, :386-412; genrec/modules/loss.py:65-77;
genrec/modules/encoder.py:399-420).  Eval mode, L2 distance, out_proj = Identity.
TEST INFRASTRUCTURE - see oracle/__init__.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, NamedTuple

import numpy as np
import torch
import torch.nn.functional as F


class RqOut(NamedTuple):
    embeddings: torch.Tensor     # [N, D, levels]
    residuals: torch.Tensor      # [N, D, levels]
    sem_ids: torch.Tensor        # [N, levels] int64
    quantize_loss: torch.Tensor  # [N]


def quantize_forward(x: torch.Tensor, codebook: torch.Tensor, commitment_weight: float = 0.25):
    """Quantize.forward, eval branch.  Follows genrec/models/rqvae.py:185-199, :246-248."""
    dist = (x ** 2).sum(1, keepdim=True) + (codebook.T ** 2).sum(0, keepdim=True) - 2 * x @ codebook.T  # :187-191
    ids = dist.min(dim=1).indices                                                # :199 (first index on ties)
    emb = codebook[ids]                                                          # :247
    loss = ((x - emb) ** 2).sum(-1) + commitment_weight * ((x - emb) ** 2).sum(-1)   # loss.py:75-77
    return emb, ids, loss


def residual_quantize(res: torch.Tensor, codebooks: List[torch.Tensor], commitment_weight: float = 0.25) -> RqOut:
    """The residual loop of RqVae.get_semantic_ids.  Follows genrec/models/rqvae.py:397-412."""
    embs, residuals, ids_all = [], [], []
    loss = 0
    for cb in codebooks:
        residuals.append(res)                                                    # :400
        emb, ids, l = quantize_forward(res, cb, commitment_weight)               # :401
        loss = loss + l                                                          # :402
        res = res - emb                                                          # :404
        ids_all.append(ids)
        embs.append(emb)
    return RqOut(torch.stack(embs, -1), torch.stack(residuals, -1), torch.stack(ids_all, -1), loss)   # :407-412


def mlp_encoder(x: torch.Tensor, weights: List[torch.Tensor]) -> torch.Tensor:
    """Bias-free Linear+SiLU stack (no activation after the last).  Follows genrec/modules/encoder.py:399-420."""
    for i, w in enumerate(weights):
        x = x @ w.T
        if i != len(weights) - 1:
            x = F.silu(x)
    return x


# --------------------------------------------------------------------------- C restatement
_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_rq.so")


def build_c(force: bool = False) -> str:
    """gcc-compile oracle/rq_argmin.c -> oracle/_build/liboracle_rq.so."""
    src = os.path.join(_HERE, "rq_argmin.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fno-fast-math", "-ffp-contract=off", "-shared", "-fPIC",
                               "-o", _SO, src, "-lm"])
    return _SO


def residual_quantize_c(res: np.ndarray, codebooks: np.ndarray):
    """Scalar C restatement.  res [N,D] f32, codebooks [levels,K,D] f32 -> ids [N,levels] i64."""
    lib = ctypes.CDLL(build_c())
    res = np.ascontiguousarray(res, dtype=np.float32)
    cbs = np.ascontiguousarray(codebooks, dtype=np.float32)
    n, d = res.shape
    lv, k, _ = cbs.shape
    ids = np.empty((n, lv), dtype=np.int64)
    out_res = np.empty((n, d), dtype=np.float32)
    lib.oracle_rq_residual_argmin(res.ctypes.data_as(ctypes.c_void_p), cbs.ctypes.data_as(ctypes.c_void_p),
                                  ctypes.c_int64(n), ctypes.c_int(d), ctypes.c_int(k), ctypes.c_int(lv),
                                  ids.ctypes.data_as(ctypes.c_void_p), out_res.ctypes.data_as(ctypes.c_void_p))
    return ids, out_res
