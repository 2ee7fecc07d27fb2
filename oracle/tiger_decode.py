"""TEST INFRASTRUCTURE - CPU restatement of the per-step post-processing of Tiger.generate (genrec/models/tiger.py:312-452).

Only tests/ may import this.  Plain Python / torch-CPU, same loops and the same dict trie as the reference; pinned by
tests/golden/tiger_decode.pt (per-step logits, multinomial draws and the final beams recorded from the UNMODIFIED reference's
``Tiger.generate`` by oracle/make_golden.py)."""
from __future__ import annotations

from collections import defaultdict

import torch


class TrieNode(defaultdict):                      # (tiger.py:40-47)
    def __init__(self):
        super().__init__(TrieNode)
        self.is_end = False


def build_trie(valid_item_ids: torch.Tensor) -> TrieNode:      # (tiger.py:49-69)
    root = TrieNode()
    flat = valid_item_ids.view(-1, valid_item_ids.size(-1)) if valid_item_ids.dim() == 3 else valid_item_ids
    for seq in flat.tolist():
        node = root
        for tok in seq:
            node = node[tok]
        node.is_end = True
    return root


DEAD_NODE = TrieNode()


def masked_log_softmax(logits: torch.Tensor, beam_nodes, B: int, K: int, vocab_offset: int, num_emb: int, temperature: float, use_trie=True):
    """(tiger.py:364-384) -> probs, log_probs [B*K, V]."""
    if use_trie:
        legal = torch.full_like(logits, False, dtype=torch.bool)
        for b in range(B):
            for k in range(K):
                valid = list(beam_nodes[b][k].keys())
                if valid:
                    legal[b * K + k, [vocab_offset + t for t in valid]] = True
        logits = logits.masked_fill(~legal, -1e32)
    else:
        mask = torch.full_like(logits, float("-inf"))
        mask[:, vocab_offset:vocab_offset + num_emb] = 0
        logits = logits + mask
    return torch.softmax(logits / temperature, dim=-1), torch.log_softmax(logits / temperature, dim=-1)


def select(beam_seqs, beam_logps, cand_token, cand_logp, beam_nodes, root, use_trie=True):
    """(tiger.py:388-441) cand_token / cand_logp [B, K, KK] (raw token ids) -> new beam_seqs [B, K, S+1], beam_logps [B, K], nodes."""
    B, K, KK = cand_token.shape
    total_logp = (beam_logps.unsqueeze(-1) + cand_logp).view(B, -1)
    total_tok = cand_token.reshape(B, -1)
    total_src = torch.arange(K).view(1, K, 1).expand(B, K, KK).reshape(B, -1)
    new_seqs, new_scores, new_nodes = [], [], []
    for b in range(B):
        scores_b, order_b = total_logp[b].sort(descending=True, stable=True)
        tokens_b, parent_b = total_tok[b][order_b], total_src[b][order_b]
        picked, seen = 0, set()
        seq = None
        for j in range(scores_b.size(0)):
            if picked == K:
                break
            p, tid = parent_b[j].item(), tokens_b[j].item()
            seq = torch.cat([beam_seqs[b, p], torch.tensor([tid])])
            key = tuple(seq.tolist())
            if key in seen:
                continue
            seen.add(key)
            picked += 1
            new_seqs.append(seq)
            new_scores.append(scores_b[j])
            if use_trie:
                new_nodes.append(beam_nodes[b][p].get(tid, DEAD_NODE))
        while picked < K:
            new_seqs.append(torch.zeros_like(seq))
            new_scores.append(torch.tensor(-1e32))
            if use_trie:
                new_nodes.append(root)
            picked += 1
    S1 = new_seqs[0].size(0)
    seqs = torch.stack(new_seqs).view(B, K, S1)
    logps = torch.stack(new_scores).view(B, K)
    nodes = [new_nodes[i * K:(i + 1) * K] for i in range(B)] if use_trie else None
    return seqs, logps, nodes


def replay(step_logits, draws, valid_item_ids, B, K, num_emb, temperature, use_trie=True):
    """Run the whole loop on recorded logits and recorded multinomial draws."""
    root = build_trie(valid_item_ids) if use_trie else None
    beam_seqs = torch.empty(B, K, 0, dtype=torch.long)
    beam_logps = torch.zeros(B, K)
    nodes = [[root for _ in range(K)] for _ in range(B)] if use_trie else None
    for step, (logits, cand) in enumerate(zip(step_logits, draws)):
        off = step * num_emb
        _, logp = masked_log_softmax(logits, nodes, B, K, off, num_emb, temperature, use_trie)
        KK = cand.size(1)
        cand_logp = torch.gather(logp, 1, cand).view(B, K, KK)
        beam_seqs, beam_logps, nodes = select(beam_seqs, beam_logps, (cand - off).view(B, K, KK), cand_logp, nodes, root, use_trie)
    return beam_seqs, beam_logps
