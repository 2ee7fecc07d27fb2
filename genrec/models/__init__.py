from genrec_b200.hstu import HSTU  # noqa: F401
from genrec_b200.rqvae import RqVae  # noqa: F401
