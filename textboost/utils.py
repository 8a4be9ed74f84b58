"""`from textboost.utils import add_augmentation_tokens, add_token, encode_prompt` (/root/reference/train_textboost.py:37-39).
`generate_prior_images` / `import_model_class_from_model_name_or_path` (class-image generation with a diffusers pipeline,
model-class dispatch on the hub config) are outside the training step and raise."""
from textboost_b200.utils import add_augmentation_tokens, add_token, encode_prompt  # noqa: F401


def import_model_class_from_model_name_or_path(pretrained_model_name_or_path: str, revision=None):
    """The reference dispatches on text_encoder/config.json `architectures`; SD checkpoints name CLIPTextModel, which the
    training script replaces with TextBoostModel anyway (/root/reference/train_textboost.py:641-651)."""
    from textboost_b200.text_encoder import CLIPTextModel
    return CLIPTextModel


def generate_prior_images(*args, **kwargs):
    raise NotImplementedError("generate_prior_images (class-image generation before training, "
                              "/root/reference/textboost/utils.py:50) is outside the training-step path; generate the "
                              "class images with inference.py and pass --class_data_dir")
