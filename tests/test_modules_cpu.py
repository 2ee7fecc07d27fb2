"""Host-side logic of the drop-in modules: state_dict schema, init, bucket tables, thresholds, error behaviour."""
import pytest
import torch

from oracle import hstu as oh


def test_hstu_state_dict_schema_matches_reference(golden):
    from genrec_b200.hstu import HSTU
    for name in ["hstu_model_d64h2.pt", "hstu_model_notime.pt"]:
        g = golden(name)
        c = g["cfg"]
        m = HSTU(c["num_items"], 64, c["embed_dim"], c["num_heads"], c["num_blocks"], use_temporal_bias=c["use_temporal_bias"])
        sd = m.state_dict()
        assert list(sd.keys()) == list(g["state_dict"].keys())
        for k, v in sd.items():
            assert v.shape == g["state_dict"][k].shape and v.dtype == g["state_dict"][k].dtype, k
        m.load_state_dict(g["state_dict"])            # strict


def test_hstu_init_statistics():
    from genrec_b200.hstu import HSTU
    torch.manual_seed(0)
    m = HSTU(2000, 50, 128, 4, 2)
    assert m.item_embedding.weight[0].abs().sum() == 0
    w = m.layers[0].projection.weight
    assert abs(w.std().item() - 0.02) < 2e-3 and w.abs().max() <= 2.0
    assert m.layers[0].projection.bias.abs().sum() == 0
    assert torch.equal(m.final_norm.weight, torch.ones(128))
    assert sum(p.numel() for p in HSTU(12101, 200, 128, 4, 4).parameters()) == 2_343_936


def test_time_thresholds_equal_reference_expression(golden):
    from genrec_b200.hstu import TemporalBias, time_bucket_thresholds
    thr = time_bucket_thresholds()
    assert thr.shape == (65,) and thr[64] == (1 << 63) - 1
    assert torch.equal(thr[:64], oh.time_bucket_thresholds(64))
    assert thr[10] == 1023 and thr[11] == 2045          # 0.693 != ln 2
    k = golden("kats.pt")
    tb = TemporalBias(64, 2)
    # the module's stand-alone forward uses the integer thresholds; check it against the reference's fp32-log buckets
    ts = torch.stack([k["dt"], torch.zeros_like(k["dt"])], 0).T.contiguous()[:4000]      # [N, 2]: dt vs 0
    bias = tb(ts)                                                                        # [N, H, 2, 2]
    want = tb.temporal_attention_bias.weight[k["dt_bucket"][:4000].long()]               # [N, H]
    assert torch.equal(bias[:, :, 0, 1], want)
    # the device formula: e = floor(log2 d); bucket = e + (d >= thr[e+1])
    d = k["dt"].clamp(min=1)
    e = torch.tensor([int(v).bit_length() - 1 for v in d.tolist()])
    b = (e + (d >= thr[e + 1]).long()).clamp(max=63)
    assert torch.equal(b.to(torch.int8), k["dt_bucket"])


def test_position_bucket_table_is_degenerate_like_the_reference(golden):
    from genrec_b200.hstu import RelativePositionBias
    pb = RelativePositionBias(32, 128, 2)
    t = pb.bucket_of_delta(300, "cpu")
    assert t.dtype == torch.uint8 and t.shape == (300,) and int(t.max()) == 0
    k = golden("kats.pt")
    L = k["rel_bucket_150"].shape[0]
    pos = torch.arange(L)
    assert torch.equal(pb._relative_position_bucket(pos[None] - pos[:, None]).to(torch.int8), k["rel_bucket_150"])
    dense = pb(L, torch.device("cpu"))
    assert dense.shape == (2, L, L)
    assert torch.equal(dense, oh.position_bias(pb.relative_attention_bias.weight, L))


def test_rqvae_mirror_schema(golden):
    from genrec_b200.rqvae import RqVae
    g = golden("rqvae_3x256x32.pt")
    c = g["cfg"]
    m = RqVae(c["input_dim"], c["D"], c["hidden_dims"], c["K"], codebook_kmeans_init=False, n_layers=c["levels"], n_cat_features=0)
    r = m.load_state_dict(g["state_dict"], strict=False)
    assert not r.unexpected_keys and all(k.startswith("decoder.") for k in r.missing_keys)
    assert m.codebooks().shape == (3, 256, 32)
    with pytest.raises(RuntimeError):
        m.eval().get_semantic_ids(g["x"])       # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        m.train().layers[0](torch.zeros(2, 32), 0.1)     # training-mode quantisation runs the CUDA search too: CPU tensors raise


def test_genrec_shim_import_paths():
    import importlib
    hstu = importlib.import_module("genrec.models.hstu")
    rq = importlib.import_module("genrec.models.rqvae")
    import genrec_b200.hstu as ours
    assert hstu.HSTU is ours.HSTU and hstu.HSTULayer is ours.HSTULayer
    assert hasattr(rq, "RqVae") and hasattr(rq, "Quantize") and hasattr(rq, "QuantizeForwardMode")
    # the two import lines of config/hstu/amazon.gin:5-6 and of config/sasrec/amazon.gin resolve with this repository alone
    dh = importlib.import_module("genrec.data.amazon_hstu")
    ds = importlib.import_module("genrec.data.amazon_sasrec")
    assert callable(dh.hstu_collate_fn) and callable(dh.hstu_eval_collate_fn) and callable(ds.sasrec_collate_fn)
    models = importlib.import_module("genrec.models")          # what genrec/models/__init__.py of the reference exports for this path
    for name in ("HSTU", "SASRec", "RqVae", "QuantizeForwardMode"):
        assert hasattr(models, name), name


def test_genrec_shim_extends_the_reference_package_instead_of_shadowing_it(tmp_path, monkeypatch):
    """With a reference checkout further down sys.path, modules this repository does not provide still resolve there."""
    import importlib
    import sys
    ref = tmp_path / "refcheckout" / "genrec"
    (ref / "trainers").mkdir(parents=True)
    (ref / "models").mkdir()
    (ref / "__init__.py").write_text("raise RuntimeError('the shim package must win')\n")
    (ref / "trainers" / "__init__.py").write_text("")
    (ref / "trainers" / "hstu_trainer.py").write_text("MARK = 'reference trainer'\n")
    (ref / "models" / "tiger.py").write_text("MARK = 'reference tiger'\n")
    monkeypatch.syspath_prepend(str(tmp_path / "refcheckout"))          # even AHEAD of us only sub-paths are merged ...
    for k in [k for k in sys.modules if k == "genrec" or k.startswith("genrec.")]:
        monkeypatch.delitem(sys.modules, k)
    import os
    monkeypatch.syspath_prepend(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # ... but this repository is first
    assert importlib.import_module("genrec.trainers.hstu_trainer").MARK == "reference trainer"
    assert importlib.import_module("genrec.models.tiger").MARK == "reference tiger"
    import genrec_b200.hstu as ours
    assert importlib.import_module("genrec.models.hstu").HSTU is ours.HSTU


def test_collate_mirrors_reference_known_answers(golden):
    from genrec_b200.data import hstu_collate_fn, hstu_eval_collate_fn, sasrec_collate_fn
    k = golden("kats.pt")
    b = [dict(history=[1, 2, 3], timestamps=[10, 20, 30], target=4), dict(history=[5], timestamps=[7], target=6)]
    got = hstu_collate_fn(b, 50)
    for key in ("input_ids", "targets", "timestamps"):
        assert torch.equal(got[key], k["hstu_collate"][key]), key
    got = hstu_eval_collate_fn(b, 2)
    for key in ("input_ids", "targets", "timestamps"):
        assert torch.equal(got[key], k["hstu_eval_collate"][key]), key
    got = sasrec_collate_fn([dict(history=[1, 2, 3], target=4), dict(history=[5], target=6)], 50)
    for key in ("input_ids", "targets"):
        assert torch.equal(got[key], k["sasrec_collate"][key]), key


def test_sasrec_state_dict_schema(golden):
    from genrec_b200.sasrec import SASRec
    g = golden("sasrec_d64h2.pt")
    c = g["cfg"]
    m = SASRec(c["num_items"], c["max_seq_len"], c["embed_dim"], c["num_heads"], c["num_blocks"], c["ffn_dim"])
    assert list(m.state_dict().keys()) == list(g["state_dict"].keys())
    m.load_state_dict(g["state_dict"])
    assert sum(p.numel() for p in SASRec(1000, 50, 64, 2, 2, 256).parameters()) == 159_040
