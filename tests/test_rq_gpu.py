import pytest
import torch

pytestmark = pytest.mark.gpu


def test_rq_vs_reference_golden(golden):
    from genrec_b200.rqvae import RqVae
    g = golden("rqvae_3x256x32.pt")
    cfg = g["cfg"]
    dev = torch.device("cuda:0")
    m = RqVae(cfg["input_dim"], cfg["D"], cfg["hidden_dims"], cfg["K"], codebook_kmeans_init=False, n_layers=cfg["levels"],
              n_cat_features=0)
    missing = m.load_state_dict(g["state_dict"], strict=False)
    assert all(k.startswith("decoder") for k in missing.missing_keys) and not missing.unexpected_keys
    m = m.to(dev).eval()
    # 1) from the reference latent: ids bit-exact, incl. the duplicated-code ties
    import genrec_b200.functional as Fn
    ids, emb, res, loss = Fn.rq_residual_argmin(g["latent"].to(dev), m.codebooks(), 0.25)
    assert torch.equal(ids.cpu(), g["sem_ids"])
    assert torch.equal(emb.cpu(), g["embeddings"])
    torch.testing.assert_close(res.cpu(), g["residuals"], rtol=0, atol=1e-6)
    torch.testing.assert_close(loss.cpu(), g["quantize_loss"], rtol=1e-5, atol=1e-6)
    # 2) end to end (encoder on cuBLAS may flip a near-tie): ids agree on >= 99% of rows
    out = m.get_semantic_ids(g["x"].to(dev))
    assert out.sem_ids.shape == g["sem_ids"].shape and out.embeddings.shape == g["embeddings"].shape
    agree = (out.sem_ids.cpu() == g["sem_ids"]).all(1).float().mean().item()
    assert agree >= 0.99, agree


@pytest.mark.parametrize("N,D,K,levels", [(12101, 32, 256, 3), (1, 32, 256, 3), (257, 64, 256, 5), (1 << 20, 32, 256, 3)])
def test_rq_vs_oracle_and_properties(N, D, K, levels):
    import genrec_b200.functional as Fn
    from oracle import rqvae as orq
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N % 1000 + D)
    x = torch.randn(N, D, generator=g)
    cbs = torch.stack([(torch.rand(K, D, generator=g) - 0.5) * (2.0 / 2 ** l) for l in range(levels)])
    ids, emb, res, loss = Fn.rq_residual_argmin(x.to(dev), cbs.to(dev), 0.25)
    ids, emb, res, loss = ids.cpu(), emb.cpu(), res.cpu(), loss.cpu()
    # properties at any size: residual recursion and optimality of every chosen code
    r = x.clone()
    for l in range(levels):
        torch.testing.assert_close(res[:, :, l], r, rtol=0, atol=1e-5)
        e = cbs[l][ids[:, l]]
        assert torch.equal(emb[:, :, l], e)
        if N <= 20000:
            d = torch.cdist(r.double(), cbs[l].double()) ** 2
            chosen = d.gather(1, ids[:, l:l + 1]).squeeze(1)
            assert (chosen <= d.min(1).values + 1e-4).all()
        r = r - e
    if N <= 20000:
        o = orq.residual_quantize(x, list(cbs))
        agree = (o.sem_ids == ids).all(1).float().mean().item()
        assert agree >= 0.995, agree
        torch.testing.assert_close(loss, o.quantize_loss, rtol=1e-4, atol=1e-5)


def test_rq_empty_and_errors():
    import genrec_b200.functional as Fn
    dev = torch.device("cuda:0")
    ids, *_ = Fn.rq_residual_argmin(torch.zeros(0, 32, device=dev), torch.rand(3, 256, 32, device=dev))
    assert ids.shape == (0, 3)
    with pytest.raises(RuntimeError):
        Fn.rq_residual_argmin(torch.zeros(4, 48, device=dev), torch.rand(3, 256, 48, device=dev))
    with pytest.raises(RuntimeError):
        Fn.rq_residual_argmin(torch.zeros(4, 32), torch.rand(3, 256, 32))
