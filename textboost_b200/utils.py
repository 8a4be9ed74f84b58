"""Host-side mirror of /root/reference/textboost/utils.py for the hot path: encode_prompt (:11-26),
add_token (:117-166) and add_augmentation_tokens (:169-214).  Same names, argument meaning and errors; the
tokenizer is duck-typed (encode / add_tokens / convert_tokens_to_ids / __len__), as no tokenizer files exist
offline."""
from __future__ import annotations


def encode_prompt(text_encoder, input_ids, attention_mask, text_encoder_use_attention_mask=None):
    """Returns the text encoder's last hidden state [B, 77, D] for `input_ids` (utils.py:11-26).

    The reference drops the mask unless the flag is set; with the flag set its mask is a Python list
    (dataset.py:428,455) and the call fails, so only the causal-mask-only path exists."""
    text_input_ids = input_ids.to(text_encoder.device)
    if text_encoder_use_attention_mask:
        attention_mask = attention_mask.to(text_encoder.device)
    else:
        attention_mask = None
    prompt_embeds = text_encoder(text_input_ids, attention_mask=attention_mask, return_dict=False)
    return prompt_embeds[0]


def add_token(text_encoder, tokenizer, placeholder_token, initializer_token):
    """Grow tokenizer + embedding matrix by one placeholder per initializer sub-token and copy the
    initializer rows into the new rows.  Returns (placeholder_tokens, placeholder_token_ids)."""
    initializer_token_ids = tokenizer.encode(initializer_token, add_special_tokens=False)
    num_vectors = len(initializer_token_ids)
    placeholder_tokens = [placeholder_token]
    if num_vectors > 1:  # multi-vector placeholder: <x> -> <x_0> <x_1> ...
        if placeholder_token.endswith(">"):
            stem = placeholder_token[:-1]
            placeholder_tokens = [f"{stem}_{i}>" for i in range(num_vectors)]
        else:
            placeholder_tokens += [f"{placeholder_token}_{i}" for i in range(1, num_vectors)]
    num_added_tokens = tokenizer.add_tokens(placeholder_tokens)
    if num_added_tokens != num_vectors:
        raise ValueError(
            f"The tokenizer already contains the token {placeholder_token}. Please pass a different"
            " `placeholder_token` that is not already in the tokenizer.")
    placeholder_token_ids = tokenizer.convert_tokens_to_ids(placeholder_tokens)
    text_encoder.resize_token_embeddings(len(tokenizer))
    token_embeds = text_encoder.get_input_embeddings().weight.data
    for token_id, initializer_token_id in zip(placeholder_token_ids, initializer_token_ids):
        token_embeds[token_id] = token_embeds[initializer_token_id].detach().clone()
    return placeholder_tokens, placeholder_token_ids


OBJECT_AUGMENTATIONS = {"<grayscale>": "grayscale", "<zoom-in>": "zoom in", "<zoom-out>": "far away",
                        "<collage>": "photo collage", "<crop>": "crop", "<hflip>": "ktn", "<left>": "pll",
                        "<right>": "ucd"}
STYLE_AUGMENTATIONS = {"<hflip>": "ktn"}


def add_augmentation_tokens(text_encoder, tokenizer, aug_type="object"):
    """Adds the augmentation pseudo-words (utils.py:169-214).  Returns (aug_token_ids, aug_token_dict)."""
    assert aug_type in ("object", "style"), f"aug_type must be either 'object' or 'style', but is {aug_type}"
    augmentations = OBJECT_AUGMENTATIONS if aug_type == "object" else STYLE_AUGMENTATIONS
    aug_token_ids, aug_token_dict = [], {}
    for placeholder_token, initializer_token in augmentations.items():
        num_vectors = len(tokenizer.encode(initializer_token, add_special_tokens=False))
        _, new_ids = add_token(text_encoder, tokenizer, placeholder_token, initializer_token)
        aug_token_ids += new_ids
        if num_vectors > 1:
            for i, token_id in enumerate(new_ids):
                aug_token_dict[placeholder_token.replace(">", f"_{i}>")] = token_id
        else:
            aug_token_dict[placeholder_token] = new_ids[0]
    return aug_token_ids, aug_token_dict
