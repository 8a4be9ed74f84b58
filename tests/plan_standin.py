"""TEST INFRASTRUCTURE ONLY: executes an ImagePlan on the CPU with the pinned oracles (oracle/pil_affine_ref.py,
oracle/pil_resample_ref.py) — the checker for the deferred augmentation: recorded plan + exact primitives must give the
bytes PIL / torchvision give when the same ops run eagerly.  The product executor is the CUDA one (image_plan.run_plan)."""
import numpy as np

from oracle import pil_affine_ref as A
from oracle import pil_resample_ref as R


def apply_op(img: np.ndarray, op: tuple) -> np.ndarray:
    kind = op[0]
    if kind == "pad_edge":
        return A.pad_edge(img, op[1], op[2])
    if kind == "affine":
        fn = A.affine_bicubic if op[7] == "bicubic" else A.affine_nearest
        return fn(img, op[1:7])
    if kind == "center_crop":
        return A.center_crop(img, op[1], op[2])
    if kind == "crop":
        x0, y0, w, h = op[1:5]
        return img[y0:y0 + h, x0:x0 + w]
    if kind == "resize":
        return R.resize(img, (op[1], op[2]), op[3])
    if kind == "flip_lr":
        return img[:, ::-1]
    if kind == "grayscale":
        return A.grayscale(img)
    if kind == "collage":
        cell = img.copy()
        cell[[0, -1], :] = 0
        cell[:, [0, -1]] = 0
        return np.tile(cell, (op[1], op[1], 1))
    raise ValueError(kind)


def run(plan) -> np.ndarray:
    from textboost_b200.image_plan import op_output_size
    img = plan.base.numpy()
    w, h = img.shape[1], img.shape[0]
    for op in plan.ops:
        img = np.ascontiguousarray(apply_op(img, op))
        w, h = op_output_size(op, w, h)
        assert (img.shape[1], img.shape[0]) == (w, h), (op, img.shape, (w, h))
    assert (img.shape[1], img.shape[0]) == plan.size
    return img
