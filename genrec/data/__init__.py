"""genrec.data: the input contract of the hot path (collate functions) from genrec_b200.data; dataset download / parsing modules
(amazon, p5_amazon, ...) resolve to a reference checkout further down sys.path."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
