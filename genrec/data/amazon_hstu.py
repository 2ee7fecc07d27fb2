"""Mirror of the collate functions of genrec/data/amazon_hstu.py:137-200 (the dataset class itself needs the Amazon dump and is
out of scope; a reference checkout on sys.path provides it through genrec.data.amazon)."""
from genrec_b200.data import hstu_collate_fn, hstu_eval_collate_fn, synthetic_batch  # noqa: F401
