"""TIGER's trie-constrained beam step on the device (csrc/beam.cuh) against the recorded decode loop of the UNMODIFIED reference
(tests/golden/tiger_decode_*.pt: per-step logits, torch.multinomial draws, final beams) and against the oracle restatement of
genrec/models/tiger.py:364-441 on random, larger cases.  Token sequences and trie nodes: bit-exact.  Log-probabilities: 1e-5."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _walk(trie, seq):
    """CSR node reached from the root along seq, -1 when the path leaves the trie."""
    off, tok, child = trie.child_off.tolist(), trie.child_tok.tolist(), trie.child_node.tolist()
    nd = 0
    for t in seq:
        nxt = -1
        for e in range(off[nd], off[nd + 1]):
            if tok[e] == t:
                nxt = child[e]
                break
        if nxt < 0:
            return -1
        nd = nxt
    return nd


@pytest.mark.parametrize("name", ["tiger_decode_trie.pt", "tiger_decode_notrie.pt"])
def test_replay_of_the_reference_decode_loop(golden, name):
    from genrec_b200 import tiger_decode as td
    g = golden(name)
    c = g["cfg"]
    dev = torch.device("cuda:0")
    trie = td.TrieCSR.build(g["valid_item_ids"]).to(dev) if c["use_trie"] else None
    logits = [x.to(dev) for x in g["step_logits"]]
    it = iter(logits)
    out = td.beam_search(lambda tgt: next(it), c["B"], c["K"], c["sem_dim"], c["num_emb"], dev, c["temperature"], trie, draws=g["draws"])
    assert torch.equal(out.sem_ids.cpu(), g["sem_ids"])
    torch.testing.assert_close(out.log_probas.cpu(), g["log_probas"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("use_trie", [True, False])
def test_masked_softmax_vs_oracle(use_trie):
    from genrec_b200 import tiger_decode as td
    from oracle import tiger_decode as od
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    B, K, num_emb, depth = 3, 5, 256, 3
    V = num_emb * depth + 1
    valid = torch.randint(0, num_emb, (400, depth), generator=g)
    root = od.build_trie(valid)
    trie = td.TrieCSR.build(valid).to(dev)
    # beams parked on assorted nodes: root, inner nodes, a leaf's parent, a dead node
    paths = [[], [valid[0, 0].item()], valid[1, :2].tolist(), [valid[2, 0].item()], [999]] * B
    nodes_o = []
    for b in range(B):
        row = []
        for k in range(K):
            nd = root
            for t in paths[b * K + k]:
                nd = nd.get(t, od.DEAD_NODE)
            row.append(nd)
        nodes_o.append(row)
    nodes_g = torch.tensor([_walk(trie, p) for p in paths[:B * K]], dtype=torch.int32, device=dev)
    for step in range(depth):
        logits = 3 * torch.randn(B * K, V, generator=g)
        po, lo = od.masked_log_softmax(logits, nodes_o, B, K, step * num_emb, num_emb, 0.2, use_trie)
        pg, lg = td.trie_log_softmax(logits.to(dev), nodes_g if use_trie else None, trie if use_trie else None, step * num_emb, num_emb, 0.2)
        torch.testing.assert_close(pg.cpu(), po, rtol=2e-5, atol=1e-7)
        fin = torch.isfinite(lo)
        assert torch.equal(torch.isfinite(lg.cpu()), fin)        # -inf outside the step's range without a trie, finite (-5e32) with one
        torch.testing.assert_close(lg.cpu()[fin], lo[fin], rtol=1e-6, atol=2e-5)
        ok = lo > -1e30                                   # the legal entries carry the information
        torch.testing.assert_close(lg.cpu()[ok], lo[ok], rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("S", [0, 1, 2])
def test_beam_select_vs_oracle(S):
    """K = 10 beams x 60 candidates (the reference's defaults, tiger.py:319,349): duplicate parents, duplicate candidates, dead and
    root nodes, fewer than K distinct sequences for one batch row."""
    from genrec_b200 import tiger_decode as td
    from oracle import tiger_decode as od
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11 + S)
    B, K, KK, num_emb = 4, 10, 60, 256
    valid = torch.randint(0, num_emb, (600, 3), generator=g)
    valid[:50, 0] = 7                                    # a crowded branch
    root = od.build_trie(valid)
    trie = td.TrieCSR.build(valid).to(dev)
    beam_seqs = valid[torch.randint(0, 600, (B, K), generator=g)][:, :, :S].contiguous()
    if S > 0:
        beam_seqs[0, 3] = beam_seqs[0, 1]                # identical parents -> identical children must be deduplicated
        beam_seqs[1, :] = beam_seqs[1, 0]                # every parent identical: fewer than K*KK distinct sequences
        beam_seqs[2, 5] = 999                            # off the trie: dead node
    beam_logps = -torch.rand(B, K, generator=g) * 3
    cand_tok = torch.stack([torch.stack([torch.randperm(num_emb, generator=g)[:KK] for _ in range(K)]) for _ in range(B)])
    cand_tok[1] = cand_tok[1, 0, :5].repeat(12)[:KK]     # row 1: only 5 distinct tokens -> 5 distinct sequences, K - 5 fillers
    cand_logp = -torch.rand(B, K, KK, generator=g) * 5
    nodes_o = []
    for b in range(B):
        row = []
        for k in range(K):
            nd = root
            for t in beam_seqs[b, k].tolist():
                nd = nd.get(t, od.DEAD_NODE)
            row.append(nd)
        nodes_o.append(row)
    nodes_g = torch.tensor([[_walk(trie, beam_seqs[b, k].tolist()) for k in range(K)] for b in range(B)], dtype=torch.int32, device=dev)
    so, lo, no = od.select(beam_seqs, beam_logps, cand_tok, cand_logp, nodes_o, root)
    sg, lg, ng = td.beam_select(beam_seqs.to(dev), beam_logps.to(dev), cand_tok.to(dev), cand_logp.to(dev), nodes_g, trie)
    assert torch.equal(sg.cpu(), so)
    assert torch.equal(lg.cpu(), lo)                      # the same fp32 addition, no re-association
    for b in range(B):
        for k in range(K):
            filler = lo[b, k].item() <= -1e31
            want = 0 if filler else _walk(trie, so[b, k].tolist())
            assert ng[b, k].item() == want, (b, k, ng[b, k].item(), want)
            # and the oracle's dict node has the same children as the CSR node
            kids = sorted(no[b][k].keys())
            if want >= 0:
                lo_, hi_ = trie.child_off[want].item(), trie.child_off[want + 1].item()
                assert trie.child_tok[lo_:hi_].tolist() == kids
            else:
                assert kids == []


def test_generate_with_a_reference_style_model():
    """End to end through generate(): a stand-in module with the reference's _encode_context / _decode_step interface."""
    from genrec_b200 import tiger_decode as td
    dev = torch.device("cuda:0")
    torch.manual_seed(0)

    class Toy(torch.nn.Module):
        sem_id_dim, num_item_embeddings = 3, 16

        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(16 * 3 + 1, 8)
            self.head = torch.nn.Linear(8, 16 * 3 + 1)

        def _encode_context(self, u, items, types, mask):
            return self.emb(items + types * 16), mask == 0

        def _decode_step(self, memory, memory_mask, tgt, types):
            h = memory.mean(1)
            if tgt is not None:
                h = h + self.emb(tgt + types * 16).sum(1)
            return self.head(h)

    m = Toy().to(dev)
    valid = torch.randint(0, 16, (60, 3))
    out = td.generate(m, torch.zeros(2, 1, dtype=torch.long, device=dev), torch.randint(0, 16, (2, 6), device=dev),
                      torch.arange(6, device=dev).remainder(3).expand(2, -1), torch.ones(2, 6, dtype=torch.long, device=dev),
                      n_top_k_candidates=5, valid_item_ids=valid)
    assert out.sem_ids.shape == (2, 5, 3) and out.log_probas.shape == (2, 5)
    items = {tuple(r) for r in valid.tolist()}
    live = out.log_probas > -1e31
    assert live.any()
    for b in range(2):
        for k in range(5):
            if live[b, k] and out.log_probas[b, k] > -1e20:
                assert tuple(out.sem_ids[b, k].tolist()) in items       # the trie only lets catalogue items through
    assert (out.log_probas[:, :-1] >= out.log_probas[:, 1:]).all()        # beams come out sorted
