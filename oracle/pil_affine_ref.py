"""oracle (test infrastructure): exact restatement of the Pillow geometry / colour arithmetic that the reference's
augmentation ops reach (/root/reference/textboost/augment/paired_augmentation.py) — groundwork for moving the
augmentation itself onto the GPU next to the byte-exact resize tail (oracle/pil_resample_ref.py):

  affine_bicubic     ``v2.functional.affine(img, ..., interpolation=BICUBIC)`` (adjust_scale, :27-31) =
                     ``Image.transform(size, AFFINE, matrix, BICUBIC)`` = Pillow ``ImagingGenericTransform`` with
                     ``affine_transform`` + ``bicubic_filter32RGB`` (src/libImaging/Geometry.c)
  affine_nearest     the same call with the default NEAREST interpolation (horizontal_translate, :117-123)
  pad_edge / center_crop   ``v2.functional.pad(..., padding_mode="edge")`` / ``center_crop`` index arithmetic
  grayscale          ``PIL.ImageOps.grayscale(img).convert("RGB")`` (:155): ITU-R 601-2 luma in 16-bit fixed point
  adjust_scale / horizontal_translate   the two composite ops, given their random draws

Third-party dependency: Pillow (12.2.0 installed here; no pin in the reference) and torchvision's
``_get_inverse_affine_matrix`` for the matrix.  PINNED against the installed libraries byte for byte by
tests/test_resample_cpu.py.  Notes on the arithmetic, all double precision: output pixel (x, y) samples the source at
(a0 (x+.5) + a1 (y+.5) + a2, ...); outside [0, W) x [0, H) the pixel is 0; otherwise the 4x4 neighbourhood around
floor(coord - .5) - 1 is combined with the a = -1 cubic (p1 + d (p2 + d (p3 + d p4))), columns clamped to the image,
rows outside the image repeating the previous row's value; the result is clamped to [0, 255] and TRUNCATED.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np


def _floor(v):
    return np.where(v < 0.0, np.floor(v), np.trunc(v)).astype(np.int64)


def _cubic(v1, v2, v3, v4, d):
    p1 = v2
    p2 = -v1 + v3
    p3 = 2 * (v1 - v2) + v3 - v4
    p4 = -v1 + v2 - v3 + v4
    return p1 + d * (p2 + d * (p3 + d * p4))


def _source_coords(H, W, matrix):
    ys, xs = np.mgrid[0:H, 0:W]
    xo, yo = xs + 0.5, ys + 0.5
    a0, a1, a2, a3, a4, a5 = matrix
    return a0 * xo + a1 * yo + a2, a3 * xo + a4 * yo + a5


def affine_bicubic(img: np.ndarray, matrix: Sequence[float]) -> np.ndarray:
    """uint8 [H, W, C] -> same shape; `matrix` = the 6 inverse-affine coefficients PIL's AFFINE transform takes."""
    H, W, _ = img.shape
    xin, yin = _source_coords(H, W, matrix)
    inside = ~((xin < 0.0) | (xin >= W) | (yin < 0.0) | (yin >= H))
    xin, yin = xin - 0.5, yin - 0.5
    x, y = _floor(xin), _floor(yin)
    dx, dy = (xin - x)[..., None], (yin - y)[..., None]
    x, y = x - 1, y - 1
    cols = [np.clip(x + i, 0, W - 1) for i in range(4)]
    f = img.astype(np.float64)

    def row(yy):
        return _cubic(f[yy, cols[0]], f[yy, cols[1]], f[yy, cols[2]], f[yy, cols[3]], dx)

    v = [row(np.clip(y, 0, H - 1))]
    for k in (1, 2, 3):
        yy = y + k
        ok = ((yy >= 0) & (yy < H))[..., None]
        v.append(np.where(ok, row(np.clip(yy, 0, H - 1)), v[-1]))
    r = _cubic(v[0], v[1], v[2], v[3], dy)
    out = np.where(r <= 0.0, 0, np.where(r >= 255.0, 255, r.astype(np.int64))).astype(np.uint8)
    out[~inside] = 0
    return out


def affine_nearest(img: np.ndarray, matrix: Sequence[float]) -> np.ndarray:
    H, W, _ = img.shape
    xin, yin = _source_coords(H, W, matrix)
    inside = ~((xin < 0.0) | (xin >= W) | (yin < 0.0) | (yin >= H))
    x = np.clip(xin.astype(np.int64), 0, W - 1)
    y = np.clip(yin.astype(np.int64), 0, H - 1)
    out = img[y, x]
    out[~inside] = 0
    return out


def pad_edge(img: np.ndarray, pad_x: int, pad_y: int) -> np.ndarray:
    return np.pad(img, ((pad_y, pad_y), (pad_x, pad_x), (0, 0)), mode="edge")


def center_crop(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """torchvision center_crop: zero-pad (left/top floor, right/bottom ceil of the deficit) any axis shorter than the
    window, then cut the window at round((size - window) / 2)."""
    H, W, _ = img.shape
    if out_h > H or out_w > W:
        px, py = max(out_w - W, 0), max(out_h - H, 0)
        img = np.pad(img, ((py // 2, (py + 1) // 2), (px // 2, (px + 1) // 2), (0, 0)))
        H, W, _ = img.shape
    top, left = int(round((H - out_h) / 2.0)), int(round((W - out_w) / 2.0))
    return img[top:top + out_h, left:left + out_w]


def grayscale(img: np.ndarray) -> np.ndarray:
    """L = (R * 19595 + G * 38470 + B * 7471 + 0x8000) >> 16, replicated to three channels."""
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    l = ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)
    return np.repeat(l[..., None], 3, axis=2)


def scale_matrix(width: int, height: int, scale: float) -> Tuple[float, ...]:
    """torchvision ``_get_inverse_affine_matrix([w/2, h/2], 0, [0, 0], scale, [0, 0])`` written out."""
    cx, cy = width * 0.5, height * 0.5
    return (1.0 / scale, 0.0, cx - cx / scale, 0.0, 1.0 / scale, cy - cy / scale)


def adjust_scale(img: np.ndarray, scale: float) -> np.ndarray:
    """The image arithmetic of paired_augmentation.adjust_scale for a given draw (incl. its (h, w) = image.size swap)."""
    H, W, _ = img.shape
    a, b = W, H  # what the reference calls (h, w)
    pad_a, pad_b = round((a / scale - a) / 2), round((b / scale - b) / 2)
    if pad_a > 0 and pad_b > 0:
        img = pad_edge(img, pad_b, pad_a)
    from torchvision.transforms.v2.functional._geometry import _get_inverse_affine_matrix
    h2, w2, _ = img.shape
    m = _get_inverse_affine_matrix([w2 * 0.5, h2 * 0.5], 0.0, [0.0, 0.0], scale, [0.0, 0.0])
    return center_crop(affine_bicubic(img, m), a, b)


def horizontal_translate(img: np.ndarray, shift: int, sign: int) -> np.ndarray:
    """paired_augmentation.horizontal_translate for a given draw: edge-pad by `shift` on both sides, move by
    sign * shift with nearest sampling, centre-crop back (output_size given as [w, h], the reference's swap)."""
    H, W, _ = img.shape
    img = pad_edge(img, shift, 0)
    from torchvision.transforms.v2.functional._geometry import _get_inverse_affine_matrix
    h2, w2, _ = img.shape
    m = _get_inverse_affine_matrix([w2 * 0.5, h2 * 0.5], 0.0, [float(sign * shift), 0.0], 1.0, [0.0, 0.0])
    return center_crop(affine_nearest(img, m), W, H)
