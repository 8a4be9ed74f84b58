"""Deferred images: the augmentation pipeline recorded as a list of primitive, exactly defined image operations instead
of being executed with PIL on the host.

`ImagePlan` quacks like the slice of ``PIL.Image.Image`` that /root/reference/textboost/augment/paired_augmentation.py
and dataset.py touch (``size`` / ``width`` / ``height``, ``transpose(FLIP_LEFT_RIGHT)``, ``crop(box)``,
``resize(size, resample)``, ``copy()``) plus the recorders the augmentation adapters call (``pad_edge``, ``affine``,
``center_crop``, ``grayscale``, ``collage``).  The ops of textboost_b200.augment run unchanged on it — same draws from
the same random streams, same caption edits — but every image operation only appends to ``ops`` and updates the size.
The list is then executed on the GPU (`run_plan`, one byte-exact kernel per primitive: csrc/augment.cu) and finished by
the resize / crop / normalise tail (image_ops).  Each primitive's arithmetic is pinned to Pillow / torchvision on the CPU
(oracle/pil_affine_ref.py, oracle/pil_resample_ref.py, tests/test_resample_cpu.py, tests/test_image_plan_cpu.py).

Primitives (all on uint8 [H, W, 3]):
  ("pad_edge", px, py)              replicate the border px columns left / right, py rows top / bottom
  ("affine", m0..m5, mode)          Pillow AFFINE transform, mode "bicubic" | "nearest", zero outside, same size
  ("center_crop", out_h, out_w)     torchvision center_crop (zero-pads an axis shorter than the window)
  ("crop", x0, y0, w, h)            window inside the image
  ("resize", w, h, filter)          Pillow antialiased resize, filter "bicubic" | "lanczos"
  ("flip_lr",)                      mirror
  ("grayscale",)                    ITU-R 601-2 luma (16-bit fixed point), replicated to RGB
  ("collage", n)                    one-pixel black frame, then an n x n tiling
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
from PIL import Image

_FILTER_NAMES = {Image.BICUBIC: "bicubic", Image.LANCZOS: "lanczos"}


class ImagePlan:
    mode = "RGB"

    def __init__(self, base: torch.Tensor, ops: Tuple[tuple, ...] = (), size: Tuple[int, int] = None, key=None):
        """key: a stable identity of the decoded base image (the file path): lets the executor keep ONE device copy
        per file however many times the plan object is re-created or pickled across DataLoader workers."""
        assert base.dtype == torch.uint8 and base.dim() == 3 and base.shape[2] == 3, "base: uint8 [H, W, 3]"
        self.base, self.ops, self.key = base, tuple(ops), key
        self._size = (int(base.shape[1]), int(base.shape[0])) if size is None else (int(size[0]), int(size[1]))

    # ---- the PIL surface the augmentation and the dataset read
    @property
    def size(self) -> Tuple[int, int]:
        return self._size

    @property
    def width(self) -> int:
        return self._size[0]

    @property
    def height(self) -> int:
        return self._size[1]

    def copy(self) -> "ImagePlan":
        return self  # immutable: every operation returns a new plan

    def _then(self, op: tuple, size: Tuple[int, int]) -> "ImagePlan":
        return ImagePlan(self.base, self.ops + (op,), size, self.key)

    def transpose(self, method) -> "ImagePlan":
        if method != Image.FLIP_LEFT_RIGHT:
            raise NotImplementedError("ImagePlan.transpose: only FLIP_LEFT_RIGHT is used by the augmentation")
        return self._then(("flip_lr",), self._size)

    def crop(self, box: Sequence[int]) -> "ImagePlan":
        x0, y0, x1, y1 = (int(v) for v in box)
        if not (0 <= x0 < x1 <= self.width and 0 <= y0 < y1 <= self.height):
            raise NotImplementedError(f"ImagePlan.crop: box {tuple(box)} leaves the {self._size} image")
        return self._then(("crop", x0, y0, x1 - x0, y1 - y0), (x1 - x0, y1 - y0))

    def resize(self, size: Sequence[int], resample=Image.BICUBIC) -> "ImagePlan":
        if resample not in _FILTER_NAMES:
            raise NotImplementedError("ImagePlan.resize: BICUBIC and LANCZOS are the filters this path uses")
        w, h = int(size[0]), int(size[1])
        if w <= 0 or h <= 0:
            raise ValueError("height and width must be > 0")
        if (w, h) == self._size:
            return self  # PIL returns a copy without resampling
        return self._then(("resize", w, h, _FILTER_NAMES[resample]), (w, h))

    # ---- recorders used by the augmentation adapters
    def pad_edge(self, pad_x: int, pad_y: int) -> "ImagePlan":
        return self._then(("pad_edge", int(pad_x), int(pad_y)), (self.width + 2 * pad_x, self.height + 2 * pad_y))

    def affine(self, matrix: Sequence[float], mode: str) -> "ImagePlan":
        assert mode in ("bicubic", "nearest") and len(matrix) == 6
        return self._then(("affine",) + tuple(float(v) for v in matrix) + (mode,), self._size)

    def center_crop(self, out_h: int, out_w: int) -> "ImagePlan":
        return self._then(("center_crop", int(out_h), int(out_w)), (int(out_w), int(out_h)))

    def grayscale(self) -> "ImagePlan":
        return self._then(("grayscale",), self._size)

    def collage(self, n: int) -> "ImagePlan":
        return self._then(("collage", int(n)), (self.width * n, self.height * n))

    def __repr__(self):
        return f"ImagePlan(base={tuple(self.base.shape)}, size={self._size}, ops={list(self.ops)})"


def op_output_size(op: tuple, width: int, height: int) -> Tuple[int, int]:
    """(width, height) after `op` on a width x height image — the size bookkeeping the executors share."""
    kind = op[0]
    if kind == "pad_edge":
        return width + 2 * op[1], height + 2 * op[2]
    if kind == "center_crop":
        return op[2], op[1]
    if kind == "crop":
        return op[3], op[4]
    if kind == "resize":
        return op[1], op[2]
    if kind == "collage":
        return width * op[1], height * op[1]
    if kind in ("affine", "flip_lr", "grayscale"):
        return width, height
    raise ValueError(f"unknown image op {kind!r}")


def gather_params(op: tuple, width: int, height: int) -> Tuple[int, int, int, int, int, int, int]:
    """(offset_x, offset_y, clamp_to_edge, flip_x, tile_w, tile_h, frame) of `tb_img_gather_u8` for the pure index
    primitives applied to a width x height image."""
    kind = op[0]
    if kind == "pad_edge":
        return -op[1], -op[2], 1, 0, 0, 0, 0
    if kind == "crop":
        return op[1], op[2], 0, 0, 0, 0, 0
    if kind == "center_crop":  # torchvision: zero-pad the short axes (floor on the left / top), cut at round(. / 2)
        out_h, out_w = op[1], op[2]
        px, py = max(out_w - width, 0), max(out_h - height, 0)
        left = int(round((width + px - out_w) / 2.0)) - px // 2
        top = int(round((height + py - out_h) / 2.0)) - py // 2
        return left, top, 0, 0, 0, 0, 0
    if kind == "flip_lr":
        return 0, 0, 0, 1, 0, 0, 0
    if kind == "collage":
        return 0, 0, 0, 0, width, height, 1
    raise ValueError(f"not an index primitive: {kind!r}")


# ------------------------------------------------------------------------------------------------ GPU executor
_device_bases = {}  # (device, plan.key) -> uint8 tensor on the device: the handful of decoded source images


def _base_on(device, plan: "ImagePlan") -> torch.Tensor:
    if plan.key is None:
        return plan.base.to(device)
    key = (str(device), plan.key, tuple(plan.base.shape))
    if key not in _device_bases:
        if len(_device_bases) >= 1024:
            _device_bases.clear()
        _device_bases[key] = plan.base.to(device)
    return _device_bases[key]


def run_plan(plan: ImagePlan, device) -> torch.Tensor:
    """Execute the recorded primitives on the GPU: uint8 [plan.height, plan.width, 3] on `device`.  One kernel per
    primitive (csrc/augment.cu; the resize is image_ops' two-pass kernel writing bytes); the decoded base image is
    uploaded once and stays resident."""
    import ctypes

    from . import _cabi as C
    from . import image_ops
    device = torch.device(device)
    img = _base_on(device, plan)
    image_ops._require_cuda(img, "run_plan")
    for op in plan.ops:
        H, W, Cc = img.shape
        w, h = op_output_size(op, W, H)
        kind = op[0]
        if kind == "resize":
            img = image_ops.resize_u8(img, (w, h), op[3])
            continue
        out = torch.empty((h, w, Cc), device=device, dtype=torch.uint8)
        if kind == "affine":
            m = (ctypes.c_double * 6)(*op[1:7])
            C.call("tb_img_affine_u8", C.ptr(img), H, W, Cc, C.ptr(out), ctypes.cast(m, ctypes.c_void_p),
                   int(op[7] == "bicubic"), C.stream_ptr())
        elif kind == "grayscale":
            C.call("tb_img_grayscale_u8", C.ptr(img), C.ptr(out), H * W, C.stream_ptr())
        else:
            C.call("tb_img_gather_u8", C.ptr(img), H, W, Cc, C.ptr(out), h, w, *gather_params(op, W, H),
                   C.stream_ptr())
        img = out
    return img
