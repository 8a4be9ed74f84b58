"""`from textboost.dataset import InstructPix2PixDataset, TextBoostDataset, PriorDataset, Wrapper`
(/root/reference/train_textboost.py:36) under the reference's class names."""
from textboost_b200.dataset import TextBoostDataset, get_images_path  # noqa: F401
from textboost_b200.prompts import HumanPromptSource as InstructPix2PixDataset  # noqa: F401
from textboost_b200.prompts import PriorPrompts as PriorDataset  # noqa: F401
from textboost_b200.prompts import ShardedStream as Wrapper  # noqa: F401
from textboost_b200.prompts import tokenize_prompt  # noqa: F401
