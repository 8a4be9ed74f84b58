"""`from textboost.augment import PairedAugmentation` (/root/reference/textboost/dataset.py)."""
from textboost_b200.augment import PairedAugmentation  # noqa: F401
