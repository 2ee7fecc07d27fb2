"""Mirror of the collate function of genrec/data/amazon_sasrec.py:125-161."""
from genrec_b200.data import sasrec_collate_fn, synthetic_batch  # noqa: F401
