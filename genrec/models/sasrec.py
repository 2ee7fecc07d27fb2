from genrec_b200.sasrec import MultiHeadAttention, PointWiseFeedForward, SASRec, SASRecBlock  # noqa: F401
