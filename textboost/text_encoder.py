"""`from textboost.text_encoder import TextBoostModel` (/root/reference/train_textboost.py:41)."""
from textboost_b200.text_encoder import CLIPTextModel, ModelOutput, TextBoostModel  # noqa: F401
