"""Prompt side of the reference's data pipeline (/root/reference/textboost/dataset.py) — the producer of the
``input_ids`` / prior ``input_ids`` the training step consumes (prompt half; the image / VAE half is
not built).  Host-only logic, mirrored so that the same seeds give the same prompts and the same rank x worker
sharding as the reference:

  tokenize_prompt              dataset.py:79-93
  TEMPLATES                    dataset.py:13-76    (public textual-inversion prompt templates)
  HumanPromptSource            dataset.py:161-193  (InstructPix2PixDataset: "input" / "output" lines of a JSONL file)
  PriorPrompts                 dataset.py:196-269  (PriorDataset: null 10 % / template 10 % / human prompt 80 %)
  ShardedStream                dataset.py:827-882  (Wrapper: shuffled, repeated, indices[rank*workers+id :: world*workers])
  instance_prompt              dataset.py:362-363  (random template formatted with the instance token)
"""
from __future__ import annotations

import json
import random
from typing import Iterator, List, Optional, Sequence

import numpy as np
import torch

_OBJ = ("a photo of a {};a rendering of a {};a cropped photo of the {};the photo of a {};a photo of a clean {};"
        "a photo of a dirty {};a dark photo of the {};a photo of my {};a photo of the cool {};a close-up photo of a {};"
        "a bright photo of the {};a cropped photo of a {};a photo of the {};a good photo of the {};a photo of one {};"
        "a close-up photo of the {};a rendition of the {};a photo of the clean {};a rendition of a {};"
        "a photo of a nice {};a good photo of a {};a photo of the nice {};a photo of the small {};"
        "a photo of the weird {};a photo of the large {};a photo of a cool {};a photo of a small {}")
_STYLE = ("a painting;a rendering;a cropped painting;the painting;a clean painting;a dirty painting;a dark painting;"
          "a picture;a cool painting;a close-up painting;a bright painting;a cropped painting;a good painting;"
          "a close-up painting;a rendition;a nice painting;a small painting;a weird painting;a large painting")
TEMPLATES = {
    "imagenet_small": _OBJ.split(";"),
    "imagenet_style_small": [s + " in the style of {}" for s in _STYLE.split(";")],
    "textboost": ["{}", "a {}", "one {}", "the {}", "photo of a {}"],
}


def resolve_template(name_or_format) -> List[str]:
    """A table name, or a literal format string used as the only template (dataset.py:292-299)."""
    return TEMPLATES.get(name_or_format, [name_or_format]) if isinstance(name_or_format, str) else [name_or_format]


def tokenize_prompt(tokenizer, prompt, tokenizer_max_length=None):
    max_length = tokenizer_max_length if tokenizer_max_length is not None else tokenizer.model_max_length
    return tokenizer(prompt, truncation=True, padding="max_length", max_length=max_length, return_tensors="pt")


def instance_prompt(template: Sequence[str], instance_token, rng=random) -> str:
    """One random template formatted with the instance token.  NB the reference passes the LIST of placeholder
    strings here (train_textboost.py:691-694), so its training prompts read "a ['<dog>']" (SURVEY.md trap 14);
    pass a string to get "a <dog>"."""
    return template[rng.randint(0, len(template) - 1)].format(instance_token)


class HumanPromptSource:
    """Prompts of a JSONL file with "input" and "output" fields; "output" is kept unless null / "NONE"."""

    def __init__(self, tokenizer, json_file, num_samples: Optional[int] = None):
        self.data: List[str] = []
        with open(json_file) as f:
            for raw in f.readlines():
                line = json.loads(raw)
                self.data.append(line["input"])
                out = line["output"]
                if out is not None and out != "NONE":
                    self.data.append(out)
        if num_samples is not None:
            self.data = self.data[:num_samples]
        self.tokenizer = tokenizer

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        prompt = self.data[index]
        t = tokenize_prompt(self.tokenizer, prompt)
        return {"prompt": prompt, "input_ids": t.input_ids, "attention_mask": t.attention_mask}


class PriorPrompts:
    """Knowledge-preservation prompts: with probability null_prob the empty prompt, with template_prob a class
    template, else the human-written prompt at `index` (one `random.random()` draw, plus `random.choice` for the
    template branch — the same draws in the same order as the reference)."""

    def __init__(self, source, tokenizer, additional_template=None, additional_category=None, template_prob=0.1,
                 null_prob=0.1):
        self.data = list(source.data)
        self.tokenizer = tokenizer
        self.template_prob, self.null_prob = template_prob, null_prob
        cats = additional_category if isinstance(additional_category, list) else [additional_category]
        self.template_data = [t.format(c) for t in resolve_template(additional_template) for c in cats]

    def __len__(self):
        return len(self.data)

    def sample_prompt(self, index, rng=random) -> str:
        r = rng.random()
        if r < self.null_prob:
            return ""
        if r < self.null_prob + self.template_prob:
            return rng.choice(self.template_data)
        return self.data[index]

    def __getitem__(self, index):
        prompt = self.sample_prompt(index)
        t = tokenize_prompt(self.tokenizer, prompt)
        return {"prompt": prompt, "input_ids": t.input_ids, "attention_mask": t.attention_mask}

    @staticmethod
    def collate_fn(samples):
        return {"prompt": [s["prompt"] for s in samples],
                "input_ids": torch.cat([s["input_ids"] for s in samples], dim=0),
                "attention_mask": [s["attention_mask"] for s in samples]}


class ShardedStream(torch.utils.data.IterableDataset):
    """Iterable over `source` sharded by rank and DataLoader worker: every epoch the index list is (optionally)
    shuffled with numpy's default_rng(seed + epoch) — cumulatively, the permutation of epoch e is applied to the
    order left by epoch e-1, as in the reference —, trimmed (drop_last) or wrapped to a multiple of world x workers,
    and this shard takes every (world x workers)-th index starting at rank x workers + worker."""

    def __init__(self, src_dataset, drop_last=True, rank: Optional[int] = None, world_size: Optional[int] = None):
        self.source, self.drop_last = src_dataset, drop_last
        self._count, self._seed, self._shuffle = 1, 0, False
        self._rank, self._world = rank, world_size

    def __len__(self):
        return len(self.source)

    def repeat(self, count=float("inf")):
        self._count = count
        return self

    def shuffle(self, mode=True, seed=None):
        if isinstance(seed, int):
            self._seed = seed
        self._shuffle = mode
        return self

    def _shard(self):
        if self._world is not None:
            world, rank = self._world, self._rank or 0
        elif torch.distributed.is_available() and torch.distributed.is_initialized():
            world, rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
        else:
            world, rank = 1, 0
        mod, shift = world, rank
        info = torch.utils.data.get_worker_info()
        if info:
            mod *= info.num_workers
            shift = shift * info.num_workers + info.id
        return mod, shift

    def indices(self) -> Iterator[int]:
        mod, shift = self._shard()
        keys = np.arange(len(self.source))
        remainder = len(keys) % mod
        epoch = 0
        while epoch < self._count:
            if self._shuffle:
                np.random.default_rng(seed=self._seed + epoch).shuffle(keys)
            if remainder == 0:
                idx = keys
            elif self.drop_last:
                idx = keys[:-remainder]
            else:
                idx = np.concatenate((keys, keys[:mod - remainder]))
            for i in idx[shift::mod]:
                yield int(i)
            epoch += 1

    def __iter__(self):
        for i in self.indices():
            yield self.source[i]
