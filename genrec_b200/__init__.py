"""genrec_b200 - B200-native (sm_100a) implementation of the sequential-attention hot path of phonism/genrec.

Public surface (mirrors the reference's import paths through the thin ``genrec`` shim package at the repo root):
    genrec_b200.hstu    HSTU, HSTULayer, RelativePositionBias, TemporalBias
    genrec_b200.sasrec  SASRec, SASRecBlock, MultiHeadAttention, PointWiseFeedForward
    genrec_b200.rqvae   Quantize, RqVae (semantic-id path)
    genrec_b200.optim   FlatAdam (fused Adam over a flat parameter buffer + bf16 mirror), DDP helper
The compute runs in ``libgenrec_b200.so`` (C ABI in include/genrec_b200.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
