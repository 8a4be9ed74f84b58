"""Import shim: the reference's package name over the B200 implementation.

`/root/reference/train_textboost.py:36-41` imports `textboost.dataset`, `textboost.utils` and `textboost.text_encoder`;
with this directory on the path those lines resolve to `textboost_b200` unchanged.  Nothing lives here:
each module re-exports the mirror of the same name."""
