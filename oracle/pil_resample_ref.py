"""oracle (test infrastructure): integer-exact restatement of Pillow's antialiased resize for 8-bit images, the
arithmetic behind ``v2.Resize(size, interpolation=LANCZOS)`` at /root/reference/textboost/dataset.py:326, 342 (and the
``Image.resize(..., BICUBIC)`` calls of the augmentation ops, paired_augmentation.py:199, 241).

The algorithm lives in a third-party dependency, Pillow (``src/libImaging/Resample.c``; version in this image: 12.2.0;
the reference pins none) — restated here from its published source and PINNED against the installed library itself:
tests/test_resample_cpu.py compares this module byte for byte with ``PIL.Image.resize`` on random images and sizes.

  * two separable passes, horizontal then vertical, each producing uint8 (the intermediate is rounded and clipped);
  * per output index: window [xmin, xmin+n) = round(center -/+ support), center = (i + 0.5) * scale,
    support = filter_support * max(scale, 1); weights filter((x + xmin - center + 0.5) / max(scale, 1)) normalised to
    sum 1 in double precision, then fixed point: int(w * 2^22 -/+ 0.5) (truncation toward zero after the signed half);
  * accumulate in int32 from 2^21 (the rounding half), shift right by 22, clip to [0, 255].
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def lanczos_filter(x: float) -> float:
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def bicubic_filter(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


FILTERS = {"lanczos": (lanczos_filter, 3.0), "bicubic": (bicubic_filter, 2.0)}


def coefficients(in_size: int, out_size: int, filter_name: str) -> Tuple[np.ndarray, np.ndarray, int]:
    """-> (bounds int32 [out, 2] = (first input index, count), kk int32 [out, ksize] fixed-point weights, ksize)."""
    fn, fsupport = FILTERS[filter_name]
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        n = xmax - xmin
        w = [fn((x + xmin - center + 0.5) * ss) for x in range(n)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, n)
    return bounds, kk, ksize


def _pass(img: np.ndarray, out_size: int, filter_name: str, axis: int) -> np.ndarray:
    """One separable pass along `axis` (0 = vertical, 1 = horizontal) of a uint8 [H, W, C] image."""
    bounds, kk, _ = coefficients(img.shape[axis], out_size, filter_name)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for i in range(out_size):
        x0, n = bounds[i]
        acc = np.tensordot(kk[i, :n].astype(np.int64), src[x0:x0 + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize(img: np.ndarray, size: Tuple[int, int], filter_name: str = "lanczos") -> np.ndarray:
    """uint8 [H, W, C] -> uint8 [size[1], size[0], C]; `size` = (width, height) as in PIL.  A pass whose size does not
    change is skipped, as in ImagingResample."""
    out_w, out_h = size
    if img.shape[1] != out_w:
        img = _pass(img, out_w, filter_name, axis=1)
    if img.shape[0] != out_h:
        img = _pass(img, out_h, filter_name, axis=0)
    return img


def shorter_side_size(width: int, height: int, size: int) -> Tuple[int, int]:
    """torchvision ``Resize(int)``: the shorter side becomes `size`, the longer int(size * long / short)."""
    short, long = (width, height) if width <= height else (height, width)
    new_short, new_long = size, int(size * long / short)
    return (new_short, new_long) if width <= height else (new_long, new_short)
