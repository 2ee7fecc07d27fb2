"""genrec.models: HSTU / SASRec / RqVae from genrec_b200 (same names the reference exports, genrec/models/__init__.py); any other
model module (tiger, lcrec, cobra, notellm) resolves to a reference checkout further down sys.path."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

from genrec_b200.hstu import HSTU  # noqa: E402,F401
from genrec_b200.rqvae import QuantizeForwardMode, RqVae  # noqa: E402,F401
from genrec_b200.sasrec import SASRec  # noqa: E402,F401
