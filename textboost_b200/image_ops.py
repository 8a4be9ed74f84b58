"""GPU tail of the image pipeline: Resize(LANCZOS) -> crop -> ToImage / ToDtype / Normalize, byte-exact.

/root/reference/textboost/dataset.py:326-351, 386-388 runs these on the host with PIL / torchvision for every item of
every step (measured here: ≈ 30 ms Lanczos 1024² -> 512² + ≈ 5 ms tensor conversion per image per core).  With
``TextBoostDataset(..., device_transforms=True)`` the dataset stops after the augmentation, hands over the uint8 image
plus the resize / crop geometry (drawn from the same random streams), and `resize_crop_normalize` finishes on the GPU
with `tb_resize_crop_normalize_u8`: Pillow's integer resampling arithmetic and torchvision's fp32 normalisation,
reproduced bit for bit (tests/test_resample_cpu.py pins the tables and the indexing to PIL on the CPU; the kernel is the
same arithmetic).  Weight tables are built on the host exactly as Pillow's precompute_coeffs / normalize_coeffs_8bpc
(double precision, sequential sums) once per (input size, output size) and cached on the device.
"""
from __future__ import annotations

import functools
import math
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

from . import _cabi as C

PRECISION_BITS = 32 - 8 - 2
_SCALE_255 = float(np.float32(1.0 / 255.0))  # torchvision: x.float().mul_(1.0 / 255)


def _lanczos(x: float) -> float:
    if not -3.0 <= x < 3.0:
        return 0.0
    if x == 0.0:
        return 1.0
    a, b = x * math.pi, x / 3 * math.pi
    return (math.sin(a) / a) * (math.sin(b) / b)


def _bicubic(x: float) -> float:
    x = abs(x)
    if x < 1.0:
        return (1.5 * x - 2.5) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * -0.5
    return 0.0


_FILTERS = {"lanczos": (_lanczos, 3.0), "bicubic": (_bicubic, 2.0)}


@functools.lru_cache(maxsize=256)
def resample_tables(in_size: int, out_size: int, filter_name: str = "lanczos"):
    """Pillow's per-axis tables: bounds int32 [out, 2] = (first tap, tap count), weights int32 [out, ksize]."""
    fn, reach = _FILTERS[filter_name]
    scale = in_size / out_size
    stretch = max(scale, 1.0)
    support = reach * stretch
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    weights = np.zeros((out_size, ksize), dtype=np.int32)
    inv = 1.0 / stretch
    one = 1 << PRECISION_BITS
    for i in range(out_size):
        center = (i + 0.5) * scale
        first = max(int(center - support + 0.5), 0)
        last = min(int(center + support + 0.5), in_size)
        taps = [fn((x + first - center + 0.5) * inv) for x in range(last - first)]
        total = 0.0
        for w in taps:  # sequential, like the C loop: the rounding of this sum decides the last weight bit
            total += w
        for x, w in enumerate(taps):
            if total != 0.0:
                w = w / total
            weights[i, x] = int(w * one - 0.5) if w < 0 else int(w * one + 0.5)
        bounds[i] = (first, last - first)
    return bounds, weights, ksize


def shorter_side_size(width: int, height: int, size: int) -> Tuple[int, int]:
    """Output (width, height) of torchvision ``Resize(int)``: shorter side = size, longer = int(size * long / short)."""
    if width <= height:
        return size, int(size * height / width)
    return int(size * width / height), size


_device_tables: Dict[tuple, tuple] = {}


def _tables_on(device, in_size, out_size, filter_name):
    key = (str(device), in_size, out_size, filter_name)
    if key not in _device_tables:
        if len(_device_tables) >= 512:  # arbitrary source sizes: bound the cache (a few KB to ~100 KB per entry)
            _device_tables.clear()
        b, w, k = resample_tables(in_size, out_size, filter_name)
        _device_tables[key] = (torch.from_numpy(b).to(device), torch.from_numpy(w).to(device), k, b)
    return _device_tables[key]


def source_rows(bounds_y: np.ndarray, top: int, crop_h: int) -> Tuple[int, int]:
    """(first source row, row count) touched by the vertical taps of resized rows [top, top + crop_h)."""
    window = bounds_y[top:top + crop_h]
    row0 = int(window[:, 0].min())
    return row0, int((window[:, 0] + window[:, 1]).max()) - row0


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what} runs on the CUDA device only (no CPU path)")


def resize_crop_normalize(src: torch.Tensor, out_size: Tuple[int, int], top: int, left: int, crop_h: int, crop_w: int,
                          out: torch.Tensor = None, filter_name: str = "lanczos", mean: float = 0.5,
                          std: float = 0.5) -> torch.Tensor:
    """src uint8 [H, W, C] on the GPU -> fp32 [C, crop_h, crop_w]: the crop window of the (out_w, out_h) resized image,
    normalised as ToDtype(float, scale=True) + Normalize(mean, std)."""
    _require_cuda(src, "resize_crop_normalize")
    assert src.dtype == torch.uint8 and src.dim() == 3 and src.is_contiguous()
    H, W, Cc = src.shape
    out_w, out_h = out_size
    bx, kx, ksx, _ = _tables_on(src.device, W, out_w, filter_name)
    by, ky, ksy, by_host = _tables_on(src.device, H, out_h, filter_name)
    row0, nrows = source_rows(by_host, top, crop_h)
    if out is None:
        out = torch.empty((Cc, crop_h, crop_w), device=src.device, dtype=torch.float32)
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (Cc, crop_h, crop_w)
    mid = torch.empty(nrows * crop_w * Cc, device=src.device, dtype=torch.uint8)
    C.call("tb_resize_crop_normalize_u8", C.ptr(src), H, W, Cc, C.ptr(bx), C.ptr(kx), ksx, out_w, C.ptr(by),
           C.ptr(ky), ksy, out_h, row0, nrows, top, left, crop_h, crop_w, _SCALE_255, float(mean), float(std),
           C.ptr(mid), C.ptr(out), None, C.stream_ptr())
    return out


def resize_u8(src: torch.Tensor, out_size: Tuple[int, int], filter_name: str = "bicubic") -> torch.Tensor:
    """Pillow's antialiased resize of a uint8 [H, W, C] image on the GPU -> uint8 [out_h, out_w, C] (the BICUBIC
    ``Image.resize`` inside the crop / collage augmentations)."""
    _require_cuda(src, "resize_u8")
    assert src.dtype == torch.uint8 and src.dim() == 3 and src.is_contiguous()
    H, W, Cc = src.shape
    out_w, out_h = out_size
    bx, kx, ksx, _ = _tables_on(src.device, W, out_w, filter_name)
    by, ky, ksy, _ = _tables_on(src.device, H, out_h, filter_name)
    out = torch.empty((out_h, out_w, Cc), device=src.device, dtype=torch.uint8)
    mid = torch.empty(H * out_w * Cc, device=src.device, dtype=torch.uint8)
    C.call("tb_resize_crop_normalize_u8", C.ptr(src), H, W, Cc, C.ptr(bx), C.ptr(kx), ksx, out_w, C.ptr(by),
           C.ptr(ky), ksy, out_h, 0, H, 0, 0, out_h, out_w, _SCALE_255, 0.5, 0.5, C.ptr(mid), None, C.ptr(out),
           C.stream_ptr())
    return out


def batch_to_pixel_values(sources: Sequence[dict], device) -> torch.Tensor:
    """``batch["sources"]`` of ``TextBoostDataset(device_transforms=True).collate_fn`` -> ``pixel_values`` fp32
    [B, 3, S, S] on `device`: one H2D copy of each augmented uint8 image, then the resize / crop / normalise kernels write
    straight into the batch."""
    S = sources[0]["crop_size"]
    batch = torch.empty((len(sources), 3, S, S), device=device, dtype=torch.float32)
    for i, s in enumerate(sources):
        if isinstance(s["source"], torch.Tensor):
            src = s["source"].to(device, non_blocking=True)
        else:  # an ImagePlan: the augmentation itself was deferred; run its primitives on the GPU first
            from .image_plan import run_plan
            src = run_plan(s["source"], device)
        top, left = s["crop_top_left"]
        resize_crop_normalize(src, tuple(s["resize_to"]), int(top), int(left), S, S, out=batch[i])
    return batch
