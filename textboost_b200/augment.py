"""Paired image / prompt augmentation of the reference's data pipeline (image half; host-only PIL
work in front of the VAE encoder).  Mirrors /root/reference/textboost/augment/paired_augmentation.py: same public
names, same argument meaning, and — because the training data a seed produces is part of the behaviour — the same
consumption of the global ``numpy.random`` / ``random`` streams, so that a seeded run draws the same crops, shifts
and caption edits as the reference (pinned by tests/golden/augment_golden.json, generated from the reference itself).

Structure here is table-driven rather than one hand-written function per caption variant: each operation is an
image transform plus a caption rule (`_Words`: the plain-text phrase, the learned-token phrase used with
``inversion=True``, and where it attaches).  Reference quirks that change the data are kept and marked "(sic)":
``image.size`` unpacked as (h, w) in the scale / crop ops, inversion phrases glued to the prompt without a space,
``<right_0>`` repeated three times.

  adjust_scale            paired_augmentation.py:20-49      crop                   :205-217
  rotate                  :52-74                            jpeg_compression       :220-233
  horizontal_flip         :77-89                            square_photo_collage   :236-260
  horizontal_translate    :92-127                           PairedAugmentation     :263-351
  adjust_brightness       :130-150                          random_resized_crop    :165-202
  grayscale               :153-162
"""
from __future__ import annotations

import math
import random
from typing import NamedTuple, Optional

import numpy as np
import PIL.Image
import PIL.ImageEnhance
import PIL.ImageOps
from PIL import Image
from torchvision.transforms import v2

from .image_plan import ImagePlan

_F = v2.functional


# ---- image primitives: executed with PIL / torchvision on a PIL image, recorded on an ImagePlan (deferred to the GPU)
def _pad_edge(image, pad_x, pad_y):
    if isinstance(image, ImagePlan):
        return image.pad_edge(pad_x, pad_y)
    return _F.pad(image, (pad_x, pad_y), padding_mode="edge")


def _affine(image, translate, scale, bicubic):
    """Scale about the centre and / or shift; torchvision's default interpolation for this call is NEAREST."""
    if isinstance(image, ImagePlan):
        from torchvision.transforms.v2.functional._geometry import _get_inverse_affine_matrix
        w, h = image.size
        matrix = _get_inverse_affine_matrix([w * 0.5, h * 0.5], 0.0, [float(translate[0]), float(translate[1])],
                                            scale, [0.0, 0.0])
        return image.affine(matrix, "bicubic" if bicubic else "nearest")
    if bicubic:
        return _F.affine(image, angle=0, translate=translate, scale=scale, shear=0, interpolation=Image.BICUBIC)
    return _F.affine(image, angle=0, translate=translate, scale=scale, shear=0)


def _center_crop(image, output_size):
    if isinstance(image, ImagePlan):
        return image.center_crop(output_size[0], output_size[1])
    return _F.center_crop(image, output_size)


def _to_grayscale(image):
    if isinstance(image, ImagePlan):
        return image.grayscale()
    return PIL.ImageOps.grayscale(image).convert("RGB")


def _framed_tiles(cell, n):
    """`cell` repeated n x n, each copy with its outermost pixel ring set to black."""
    if isinstance(cell, ImagePlan):
        return cell.collage(n)
    px = np.asarray(cell).copy()
    px[[0, -1], :] = 0
    px[:, [0, -1]] = 0
    return Image.fromarray(np.tile(px, (n, n, 1)))


def _pil_only(image, what):
    if isinstance(image, ImagePlan):
        raise NotImplementedError(f"{what} is not part of the deferred (GPU) augmentation path: PairedAugmentation "
                                  "never schedules it")


class _Words(NamedTuple):
    plain: str   # caption words without learned tokens
    tokens: str  # augmentation-token phrase (inversion=True); the tokens are registered by utils.add_augmentation_tokens

    def pick(self, inversion: bool) -> str:
        return self.tokens if inversion else self.plain


def _front(prompt: str, words: str, sep: str = "") -> str:
    return words + sep + prompt


def _back(prompt: str, words: str, sep: str = ", ") -> str:
    return prompt + sep + words


def _either_end(prompt: str, words: str, front_sep: str) -> str:
    """One uniform draw: below one half the words go in front, otherwise behind a comma."""
    return _front(prompt, words, front_sep) if np.random.random() < 0.5 else _back(prompt, words)


# ---------------------------------------------------------------------------------------------- geometric operations
def adjust_scale(image, prompt, inversion=False):
    """Zoom by a factor drawn from U[0.34, 1.4] about the centre, edge-padded when zooming out."""
    s = np.random.uniform(0.34, 1.4)
    a, b = image.size  # (sic) the reference names these (h, w); PIL gives (width, height)
    pad_a, pad_b = round((a / s - a) / 2), round((b / s - b) / 2)
    if pad_a > 0 and pad_b > 0:
        image = _pad_edge(image, pad_b, pad_a)
    image = _affine(image, (0, 0), s, bicubic=True)
    image = _center_crop(image, (a, b))
    if inversion:
        words = "<zoom-out_0> <zoom-out_1>" if s < 0.6 else "<zoom-in_0> <zoom-in_1>" if s > 1.2 else ""
    elif s <= 0.6:
        words = ("a far away ", "very small ")[np.random.choice(2)]
    elif s >= 1.2:
        words = ("zoomed in ", "close up ")[np.random.choice(2)]
    else:
        words = ""
    return image, _front(prompt, words)


_ROTATIONS = ((90, _Words("90 degrees counter clockwise rotated ", "<rot90_0> <rot90_1>")),
              (-90, _Words("90 degrees clockwise rotated ", "<rot270_0> <rot270_1>")))


def rotate(image, prompt, inversion=False):
    angle, words = _ROTATIONS[np.random.randint(0, 2)]
    _pil_only(image, "rotate")
    image = _F.rotate(image, angle=angle)
    if inversion:
        return image, _either_end(prompt, words.tokens, "")
    return image, _front(prompt, words.plain)


def horizontal_flip(image, prompt, inversion=False):
    image = image.transpose(Image.FLIP_LEFT_RIGHT)
    if inversion:
        return image, _either_end(prompt, "<hflip>", " ")
    return image, _either_end(prompt, "horizontally flipped", " ")


_SHIFTS = ((-1, _Words(" on the left", " <left_0> <left_1> <left_2>")),
           (+1, _Words(" on the right", " <right_0> <right_0> <right_0>")))  # (sic) right_0 three times


def horizontal_translate(image, prompt, inversion=False):
    """Shift by 15-30 % of the width to either side; the uncovered strip is filled from the edge column."""
    sign, words = _SHIFTS[np.random.randint(0, 2)]
    w, h = image.size
    shift = int(np.random.uniform(low=0.15, high=0.3) * w)
    image = _pad_edge(image, shift, 0)
    image = _affine(image, [sign * shift, 0], 1, bicubic=False)
    image = _center_crop(image, [w, h])  # (sic) output_size is (height, width); equal for the square inputs used
    return image, _back(prompt, words.pick(inversion), sep="")


def random_resized_crop(image, target_size, scale=(0.08, 1.0), ratio=(3. / 4., 4. / 3.)):
    """A window of random area (fraction of the image in `scale`) and aspect ratio (in `ratio`) at a random position,
    resized to `target_size` = (width, height) with bicubic filtering.  Draws come from the ``random`` module."""
    width, height = image.size
    area = width * height * random.uniform(*scale)
    aspect = random.uniform(*ratio)
    win_w = min(int(round(math.sqrt(area * aspect))), width)
    win_h = min(int(round(math.sqrt(area / aspect))), height)
    x0 = random.randint(0, width - win_w)
    y0 = random.randint(0, height - win_h)
    return image.crop((x0, y0, x0 + win_w, y0 + win_h)).resize(target_size, Image.BICUBIC)


def crop(image, prompt, inversion=False):
    a, b = image.size  # (sic) as in adjust_scale
    image = random_resized_crop(image, (a, b), ratio=(1.0, 1.0))
    return image, _either_end(prompt, _Words("cropped", "<crop>").pick(inversion), " ")


# -------------------------------------------------------------------------------------------------- colour operations
def adjust_brightness(image, prompt, inversion=False, size=None):
    del size
    if np.random.random() < 0.5:
        factor, words = np.random.uniform(0.4, 0.6), _Words("dimmed", "<dimmed>")
    else:
        factor, words = np.random.uniform(1.3, 1.5), _Words("bright", "<bright>")
    _pil_only(image, "adjust_brightness")
    image = PIL.ImageEnhance.Brightness(image).enhance(factor)
    return image, _either_end(prompt, words.pick(inversion), "")


def grayscale(image, prompt, inversion=False, size=None):
    del size
    image = _to_grayscale(image)
    return image, _back(prompt, _Words("grayscale", "<grayscale_0> <grayscale_1>").pick(inversion))


def jpeg_compression(image, prompt, inversion=False):
    quality = np.random.randint(25, 75)
    _pil_only(image, "jpeg_compression")
    image = _F.jpeg(image, quality=quality)
    return image, _either_end(prompt, _Words("JPEG", "<jpeg_0> <jpeg_1>").pick(inversion), " ")


# --------------------------------------------------------------------------------------------------- other operations
def square_photo_collage(image, prompt, inversion=False):
    """The image shrunk to a 2x2 or 3x3 grid of copies, each with a one-pixel black frame."""
    n = np.random.randint(2, 4)
    w, h = image.size
    cell_w, cell_h = w // n, h // n
    image = _framed_tiles(image.resize((cell_h, cell_w), Image.BICUBIC), n)  # PIL size (cell_h, cell_w): (sic)
    return image, _front(prompt, _Words("photo collage of ", "<collage_0> <collage_1> ").pick(inversion))


# ------------------------------------------------------------------------------------------------------- the pipeline
_OBJECT_STAGES = ((adjust_scale, crop, horizontal_translate), (square_photo_collage,), (grayscale,))
_STYLE_STAGES = ((), (), (grayscale,))


class PairedAugmentation:
    """Three stages — geometric, other, colour — each applied with its own probability (``p``, ``p``,
    ``color_prob``), one operation drawn uniformly per applied stage; the prompt follows the image only when
    ``augment_prompt`` is set.  ``hflip``: "true" = silent random flip, "inversion" = flip as a captioned geometric
    operation, "false" = none.  Returns (image, prompt, None) — the mask slot is unused on this path."""

    def __init__(self, hflip="false", inversion=False, p=0.5, color_prob=0.2, augment_prompt=True, ops="object"):
        mode = hflip.lower()
        assert mode in ("true", "false", "inversion"), f"Invalid hflip value: {hflip}"
        self.hflip = mode == "true"
        self.inversion = inversion
        self.p = p
        self.color_prob = color_prob
        self.augment_prompt = augment_prompt
        geometric, other, color = _OBJECT_STAGES if ops == "object" else _STYLE_STAGES
        self.geometric_ops = list(geometric)
        self.color_ops = list(color)
        self.other_ops = list(other)
        if mode == "inversion":
            self.geometric_ops.append(horizontal_flip)

    def _stage(self, ops, prob, image, prompt):
        # the probability draw is skipped for an empty stage, the operation index is one np.random.choice draw
        if not ops or not np.random.rand() < prob:
            return image, prompt
        image, captioned = ops[np.random.choice(len(ops))](image, prompt, self.inversion)
        return image, (captioned if self.augment_prompt else prompt)

    def __call__(self, image, prompt):
        assert isinstance(image, (PIL.Image.Image, ImagePlan)), \
            f"Invalid image type ({type(image)}). Must be PIL.Image.Image."
        if self.hflip and np.random.rand() < 0.5:
            image = image.transpose(Image.FLIP_LEFT_RIGHT)
        image, prompt = self._stage(self.geometric_ops, self.p, image, prompt)
        image, prompt = self._stage(self.other_ops, self.p, image, prompt)
        image, prompt = self._stage(self.color_ops, self.color_prob, image, prompt)
        mask: Optional[np.ndarray] = None
        return image, prompt, mask
