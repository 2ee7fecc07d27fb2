from genrec_b200.hstu import HSTU, HSTULayer, RelativePositionBias, TemporalBias  # noqa: F401
