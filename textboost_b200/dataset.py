"""Image side of the reference's data pipeline: instance images -> augmented, resized, cropped,
normalised ``pixel_values`` + tokenised prompts, batched the way the training loop consumes them
(/root/reference/train_textboost.py:872-890, 1027-1037).  Host-only PIL / torchvision work; the pixels then go to
``textboost_b200.vae`` on the GPU.  Mirrors /root/reference/textboost/dataset.py with the same names and draws:

  get_images_path          dataset.py:96-105
  TextBoostDataset         dataset.py:272-418   (item dict keys: image, input_ids, attention_mask, original_size,
                                                 crop_top_left, mask / class_* when present)
  TextBoostDataset.collate_fn   dataset.py:420-459
  InstructPix2PixDataset / PriorDataset / Wrapper  -> textboost_b200.prompts (re-exported under the reference names)

Random draws per item, in order: ``random.randint`` for the template, the augmentation pipeline's numpy / random
draws, torch's RNG for the random crop — identical to the reference so that a seed reproduces its training data
(tests/golden/dataset_golden.json is generated from the reference classes).
"""
from __future__ import annotations

import os
import random
from pathlib import Path
from typing import List, Optional, Sequence

import torch
from PIL import Image
from PIL.ImageOps import exif_transpose
from torchvision.transforms import v2

from .prompts import HumanPromptSource as InstructPix2PixDataset  # noqa: F401  (reference import names)
from .prompts import PriorPrompts as PriorDataset  # noqa: F401
from .prompts import ShardedStream as Wrapper  # noqa: F401
from .prompts import resolve_template, tokenize_prompt


def get_images_path(data_root, max_samples: Optional[int] = None) -> List[Path]:
    """Sorted directory listing, optionally truncated (``--num_samples``)."""
    root = Path(data_root)
    if not root.exists():
        raise ValueError("Data root doesn't exists.")
    paths = sorted(root.iterdir())
    return paths if max_samples is None else paths[:max_samples]


def _open_rgb(path) -> Image.Image:
    image = exif_transpose(Image.open(path))
    return image if image.mode == "RGB" else image.convert("RGB")


class TextBoostDataset(torch.utils.data.Dataset):
    def __init__(self, concepts_list: Sequence[dict], tokenizer, tokenizer_2=None, num_instance=None, template="a {}",
                 prior_data_root=None, class_token=None, num_prior=None, size=512, center_crop=False,
                 augment_pipe=None, augment_prior: bool = False, device_transforms: bool = False,
                 cache_decoded: bool = False, device_augment: bool = False):
        """device_transforms (addition): stop after the augmentation and return the uint8 image ("source" [H,W,3]) with
        the resize / crop geometry ("resize_to", "crop_top_left", "crop_size") instead of "image"; the byte-exact GPU
        tail (textboost_b200.image_ops) produces the same ``pixel_values``.  The random streams are consumed exactly
        as without it.  cache_decoded (addition): keep the decoded RGB source images (a handful in this workload)
        instead of re-opening the files for every item; the pixels are the same.  device_augment (addition, implies
        the other two): instance images become `ImagePlan`s — the augmentation pipeline draws its random numbers and
        edits the caption as always but only RECORDS its image operations, which then run as exact GPU kernels
        (textboost_b200.image_plan.run_plan); "source" is the plan."""
        self.size, self.center_crop = size, center_crop
        self.device_augment = device_augment
        device_transforms, cache_decoded = device_transforms or device_augment, cache_decoded or device_augment
        self.device_transforms, self._decoded = device_transforms, ({} if cache_decoded else None)
        self._bases = {}
        self.tokenizer, self.tokenizer_2 = tokenizer, tokenizer_2
        self.template = resolve_template(template)
        self.instance_images_path = [(path, concept["instance_token"]) for concept in concepts_list
                                     for path in get_images_path(concept["instance_data_dir"], num_instance)]
        self.num_instance_images = self._length = len(self.instance_images_path)
        self.class_token = class_token
        self.prior_data_root = None
        if prior_data_root is not None:
            self.prior_data_root = Path(prior_data_root)
            self.prior_data_root.mkdir(parents=True, exist_ok=True)
            self.class_images_path = list(self.prior_data_root.iterdir())
            n = len(self.class_images_path)
            self.num_prior_images = n if num_prior is None else min(n, num_prior)
            self._length = max(self.num_prior_images, self.num_instance_images)
        self.resize_fn = v2.Resize(size, interpolation=v2.InterpolationMode.LANCZOS)
        self.crop = v2.CenterCrop(size) if center_crop else v2.RandomCrop(size)
        # PIL -> float CHW in [-1, 1]
        self.image_transforms = v2.Compose([v2.ToImage(), v2.ToDtype(torch.float, scale=True),
                                            v2.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
        self.augment_pipe, self.augment_prior = augment_pipe, augment_prior

    def __len__(self):
        return self._length

    def _open(self, path):
        if self._decoded is None:
            return _open_rgb(path)
        if path not in self._decoded:
            self._decoded[path] = _open_rgb(path)
            self._decoded[path].load()
        return self._decoded[path].copy()

    def _open_plan(self, path):
        """The decoded image as the base of a deferred ImagePlan (one uint8 tensor per file, shared by every item)."""
        import numpy as np
        from .image_plan import ImagePlan
        if path not in self._bases:
            self._bases[path] = torch.from_numpy(np.array(_open_rgb(path), dtype=np.uint8))
        return ImagePlan(self._bases[path], key=str(path))

    def _geometry(self, image):
        """What `_resize_and_crop_image` would do to `image`, without doing it: (resized (w, h), top, left), drawing the
        random crop position from torch's RNG exactly as RandomCrop.get_params does on the resized image."""
        from .image_ops import shorter_side_size
        w, h = shorter_side_size(image.width, image.height, self.size)
        if self.center_crop:
            return (w, h), max(0, int(round((h - self.size) / 2.0))), max(0, int(round((w - self.size) / 2.0)))
        top, left, _, _ = self.crop.get_params(torch.empty((3, h, w), dtype=torch.uint8), (self.size, self.size))
        return (w, h), top, left

    def _finish(self, sample, image, prefix=""):
        """Resize / crop / normalise on the host (reference behaviour) or record the geometry for the GPU tail."""
        if not self.device_transforms:
            image, top, left = self._resize_and_crop_image(image)
            sample["class_image" if prefix else "image"] = self.image_transforms(image)
            sample[prefix + "crop_top_left"] = (top, left)
            return
        import numpy as np
        resize_to, top, left = self._geometry(image)
        from .image_plan import ImagePlan
        if isinstance(image, ImagePlan):
            sample[prefix + "source"] = image
        else:
            sample[prefix + "source"] = torch.from_numpy(np.array(image, dtype=np.uint8))  # [H, W, 3], writable copy
        sample[prefix + "resize_to"], sample[prefix + "crop_size"] = resize_to, self.size
        sample[prefix + "crop_top_left"] = (top, left)

    def _resize_and_crop_image(self, image):
        """Shorter side to `size` (Lanczos), then a `size` x `size` window: centred, or at a position drawn from
        torch's RNG.  Returns the window and its top-left corner (y, x)."""
        image = self.resize_fn(image)
        if self.center_crop:
            top = max(0, int(round((image.height - self.size) / 2.0)))
            left = max(0, int(round((image.width - self.size) / 2.0)))
            return self.crop(image), top, left
        top, left, h, w = self.crop.get_params(image, (self.size, self.size))
        return v2.functional.crop(image, top, left, h, w), top, left

    def _augment(self, image, prompt, sample, mask_key):
        image, prompt, mask = self.augment_pipe(image, prompt)
        if mask is not None:
            sample[mask_key] = torch.as_tensor(mask, dtype=torch.float32).unsqueeze(0)
        return image, prompt

    def _tokenize_into(self, sample, prompt, prefix=""):
        t = tokenize_prompt(self.tokenizer, prompt)
        sample[prefix + "input_ids"], sample[prefix + "attention_mask"] = t.input_ids, t.attention_mask
        if self.tokenizer_2 is not None and not prefix:
            t2 = tokenize_prompt(self.tokenizer_2, prompt)
            sample["input_ids_2"], sample["attention_mask_2"] = t2.input_ids, t2.attention_mask

    def __getitem__(self, index):
        sample = {}
        path, instance_token = self.instance_images_path[index % self.num_instance_images]
        image = self._open_plan(path) if self.device_augment else self._open(path)
        which = random.randint(0, len(self.template) - 1)
        prompt = self.template[which].format(instance_token)
        if self.augment_pipe is not None:
            image, prompt = self._augment(image, prompt, sample, "mask")
        sample["original_size"] = (image.width, image.height)
        self._finish(sample, image)
        self._tokenize_into(sample, prompt)
        if self.prior_data_root:
            self._class_item(sample, index, which)
        return sample

    def _class_item(self, sample, index, which):
        """The ``--with_image_prior`` half of an item: a class image with the same template, or with the prompt
        spelled in its file name (``<n>-<words_with_underscores>.<ext>``) when no class token is given."""
        path = self.class_images_path[index % self.num_prior_images]
        image = self._open_plan(path) if self.device_augment else exif_transpose(Image.open(path)).convert("RGB")
        if self.class_token is not None:
            prompt = self.template[which].format(self.class_token)
        else:
            prompt = os.path.basename(path).split("-")[1].split(".")[0].replace("_", " ")
        if self.augment_prior and self.augment_pipe is not None:
            image, prompt = self._augment(image, prompt, sample, "prior_mask")
        if "mask" in sample and "prior_mask" not in sample:
            sample["prior_mask"] = torch.ones_like(sample["mask"])
        # the reference resizes and crops once, then runs the resize-and-crop helper on the result: with a random
        # crop that is two draws from torch's RNG, kept so that seeded runs stay aligned
        image = self._first_class_crop(image)
        self._finish(sample, image, prefix="class_")
        self._tokenize_into(sample, prompt, prefix="class_")

    def _first_class_crop(self, image):
        """``self.crop(self.resize_fn(image))`` (dataset.py:408-409), eagerly on a PIL image or recorded on a plan.  The
        random crop of this call is the transform's own forward: it draws only for the axes that need cropping (unlike
        `get_params`, which `_resize_and_crop_image` uses)."""
        from .image_plan import ImagePlan
        if not isinstance(image, ImagePlan):
            return self.crop(self.resize_fn(image))
        from .image_ops import shorter_side_size
        w, h = shorter_side_size(image.width, image.height, self.size)
        image = image.resize((w, h), Image.LANCZOS)
        if self.center_crop:
            top, left = int(round((h - self.size) / 2.0)), int(round((w - self.size) / 2.0))
        else:
            params = self.crop.make_params([torch.empty((3, h, w), dtype=torch.uint8)])
            assert not params["needs_pad"]
            top, left = params["top"], params["left"]
        return image.crop((left, top, left + self.size, top + self.size))

    @staticmethod
    def collate_fn(samples, with_prior_preservation=False):
        """-> {"input_ids" [n,L], "pixel_values" [n,3,S,S] fp32, "attention_mask" list, "mask" if present}; with
        prior preservation the class examples are appended after the instance ones (one forward for both)."""
        keys = [""] + (["class_"] if with_prior_preservation else [])
        has_mask = "attention_mask" in samples[0]
        ids = [s[k + "input_ids"] for k in keys for s in samples]
        batch = {"input_ids": torch.cat(ids, dim=0)}
        if "source" in samples[0]:  # device_transforms: uint8 sources + geometry, finished by image_ops on the GPU
            batch["sources"] = [{n: s[k + n] for n in ("source", "resize_to", "crop_top_left", "crop_size")}
                                for k in keys for s in samples]
        else:
            pixels = [s["class_image" if k else "image"] for k in keys for s in samples]
            batch["pixel_values"] = torch.stack(pixels).to(memory_format=torch.contiguous_format).float()
        if "mask" in samples[0]:
            masks = [s["mask"] for s in samples]
            if "prior_mask" in samples[0]:
                masks += [s["prior_mask"] for s in samples]
            batch["mask"] = torch.stack(masks)
        if has_mask:
            batch["attention_mask"] = [s[k + "attention_mask"] for k in keys for s in samples]
        return batch
