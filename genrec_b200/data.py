"""Input contract of the hot path: collate functions (mirrors of genrec/data/amazon_hstu.py:137-200 and
genrec/data/amazon_sasrec.py:125-181), synthetic generators and data-parallel batch sharding.
Host-side, pure Python/torch-CPU; nothing here computes model arithmetic."""
from __future__ import annotations

from typing import Dict, List

import torch


def hstu_collate_fn(batch: List[Dict], max_seq_len: int = 50):
    """LEFT-pad to the batch maximum (not max_seq_len), targets shifted by one, padded timestamps = 0
    (genrec/data/amazon_hstu.py:137-173)."""
    histories = [b["history"] for b in batch]
    stamps = [b["timestamps"] for b in batch]
    targets = [b["target"] for b in batch]
    max_len = min(max(len(h) for h in histories), max_seq_len)
    ids, tgs, tss = [], [], []
    for h, ts, t in zip(histories, stamps, targets):
        if len(h) > max_len:
            h, ts = h[-max_len:], ts[-max_len:]
        seq = list(h) + [t]
        ts_seq = list(ts) + [ts[-1] if ts else 0]
        pad = max_len + 1 - len(seq)
        seq, ts_seq = [0] * pad + seq, [0] * pad + ts_seq
        ids.append(seq[:-1]); tgs.append(seq[1:]); tss.append(ts_seq[:-1])
    return {"input_ids": torch.tensor(ids, dtype=torch.long), "targets": torch.tensor(tgs, dtype=torch.long),
            "timestamps": torch.tensor(tss, dtype=torch.long)}


def hstu_eval_collate_fn(batch: List[Dict], max_seq_len: int = 50):
    """genrec/data/amazon_hstu.py:176-200: targets is the single next item per sample."""
    histories = [b["history"] for b in batch]
    stamps = [b["timestamps"] for b in batch]
    max_len = min(max(len(h) for h in histories), max_seq_len)
    ids, tss = [], []
    for h, ts in zip(histories, stamps):
        if len(h) > max_len:
            h, ts = h[-max_len:], ts[-max_len:]
        pad = max_len - len(h)
        ids.append([0] * pad + list(h)); tss.append([0] * pad + list(ts))
    return {"input_ids": torch.tensor(ids, dtype=torch.long), "targets": torch.tensor([b["target"] for b in batch], dtype=torch.long),
            "timestamps": torch.tensor(tss, dtype=torch.long)}


def sasrec_collate_fn(batch: List[Dict], max_seq_len: int = 50):
    """genrec/data/amazon_sasrec.py:125-161 (same as the HSTU one without timestamps)."""
    out = hstu_collate_fn([dict(history=b["history"], timestamps=[0] * len(b["history"]), target=b["target"]) for b in batch], max_seq_len)
    return {"input_ids": out["input_ids"], "targets": out["targets"]}


def synthetic_batch(B: int, L: int, V: int, seed: int, full_length: bool = True):
    """ids ~ Zipf(1.1) over 1..V, timestamps = 1.30e9 + cumsum(Exp(mean 3 days)); with
    full_length=False the lengths are ~U[L/4, L] and the batch is left-padded exactly like hstu_collate_fn."""
    g = torch.Generator().manual_seed(seed)
    w = torch.arange(1, V + 1, dtype=torch.float64).pow(-1.1)
    ids = torch.multinomial(w, B * (L + 1), replacement=True, generator=g).view(B, L + 1) + 1
    gaps = torch.empty(B, L).exponential_(1.0 / (3 * 86400.0), generator=g).long() + 1
    ts = 1_300_000_000 + torch.cumsum(gaps, 1)
    inp, tgt = ids[:, :L].clone(), ids[:, 1:].clone()
    if not full_length:
        lens = torch.randint(max(1, L // 4), L + 1, (B,), generator=g)
        for b in range(B):
            p = L - int(lens[b])
            inp[b, :p] = 0; ts[b, :p] = 0
            tgt[b, :max(p - 1, 0)] = 0
    return inp.contiguous(), ts.contiguous(), tgt.contiguous()


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Contiguous equal split of the global batch across data-parallel ranks (what Accelerate's prepared DataLoader does with
    split_batches=True; with the default split_batches=False every rank simply draws its own batch)."""
    out = {}
    for k, v in batch.items():
        n = v.shape[0]
        assert n % world == 0, "global batch must divide evenly across ranks"
        per = n // world
        out[k] = v[rank * per:(rank + 1) * per]
    return out


def collate_jagged(items: torch.Tensor, offsets: torch.Tensor, targets: torch.Tensor, max_seq_len: int = 50,
                   timestamps: "torch.Tensor | None" = None, max_len_in_batch: "int | None" = None) -> Dict[str, torch.Tensor]:
    """hstu_collate_fn / sasrec_collate_fn ON THE DEVICE: a jagged batch that already lives in HBM (items / timestamps [N] int64 in time
    order, offsets [B+1], one held-out target per user) -> the left-padded [B, L] batch dict, without a host round trip.
    L = min(longest history, max_seq_len); pass ``max_len_in_batch`` (the loader knows it) to avoid the one device sync that reading it
    from ``offsets`` costs."""
    from . import _lib
    from ._lib import check, ptr, require_cuda, stream_ptr
    require_cuda(items, offsets, targets)
    for t in (items, offsets, targets, timestamps):
        if t is not None and t.dtype != torch.int64:
            raise _lib.GrbError(f"genrec_b200 error -1: jagged batches are int64 (got {t.dtype})")
    B = offsets.numel() - 1
    if max_len_in_batch is None:
        max_len_in_batch = int((offsets[1:] - offsets[:-1]).max().item())
    L = max(1, min(int(max_len_in_batch), int(max_seq_len)))
    dev = items.device
    ids = torch.empty(B, L, dtype=torch.int64, device=dev)
    tgs = torch.empty(B, L, dtype=torch.int64, device=dev)
    tss = torch.empty(B, L, dtype=torch.int64, device=dev) if timestamps is not None else None
    with torch.cuda.device(dev):
        check(_lib.load().grb_collate_jagged(ptr(items.contiguous()), ptr(timestamps.contiguous()) if timestamps is not None else None,
                                             ptr(offsets.contiguous()), ptr(targets.contiguous()), B, L, ptr(ids), ptr(tgs), ptr(tss),
                                             stream_ptr(dev)))
    out = {"input_ids": ids, "targets": tgs}
    if tss is not None:
        out["timestamps"] = tss
    return out
