"""TEST INFRASTRUCTURE ONLY: a fake `_cabi.call` for the image entry points that reads and writes the raw HOST pointers
it is handed, with numpy transcriptions of the kernels' index arithmetic.  It lets the CPU suite check the HOST side of
the image path — argument order and meaning of the 25-argument C calls, buffer sizes, strides, table upload, op-to-
parameter mapping — end to end against PIL; the CUDA kernels themselves are checked on the GPU.  `install(monkeypatch)`
also lifts the "CUDA only" guards for the duration of one test; the product path has no CPU route."""
import ctypes

import numpy as np


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, ctypes.c_void_p):
        return p.value or 0
    return int(p)


def _arr(p, shape, dtype):
    a = _addr(p)
    assert a, "null pointer"
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    return np.frombuffer((ctypes.c_char * n).from_address(a), dtype=dtype).reshape(shape)


def tb_resize_crop_normalize_u8(src, H, W, C, bx, kx, ksx, out_w, by, ky, ksy, out_h, row0, nrows, top, left, ch, cw,
                                scale, mean, std, mid, out_f32, out_u8, stream):
    src = _arr(src, (H, W, C), np.uint8)
    bx, kx = _arr(bx, (out_w, 2), np.int32), _arr(kx, (out_w, ksx), np.int32)
    by, ky = _arr(by, (out_h, 2), np.int32), _arr(ky, (out_h, ksy), np.int32)
    assert 0 <= top and top + ch <= out_h and 0 <= left and left + cw <= out_w and row0 + nrows <= H
    mid = _arr(mid, (nrows, cw, C), np.uint8)
    half = np.int32(1 << 21)
    for j in range(cw):
        x0, n = bx[left + j]
        acc = (src[row0:row0 + nrows, x0:x0 + n].astype(np.int32) * kx[left + j, :n, None]).sum(1, dtype=np.int32) + half
        mid[:, j] = np.clip(acc >> 22, 0, 255)
    f32 = _arr(out_f32, (C, ch, cw), np.float32) if _addr(out_f32) else None
    u8 = _arr(out_u8, (ch, cw, C), np.uint8) if _addr(out_u8) else None
    for r in range(ch):
        y0, n = by[top + r]
        acc = (mid[y0 - row0:y0 - row0 + n].astype(np.int32) * ky[top + r, :n, None, None]).sum(0, dtype=np.int32) + half
        v = np.clip(acc >> 22, 0, 255)
        if u8 is not None:
            u8[r] = v
        if f32 is not None:
            f = v.astype(np.float32) * np.float32(scale)
            f32[:, r] = ((f - np.float32(mean)) / np.float32(std)).T
    return 0


def tb_img_gather_u8(src, sh, sw, C, out, oh, ow, ox, oy, clamp, flip, tw, th, frame, stream):
    src, out = _arr(src, (sh, sw, C), np.uint8), _arr(out, (oh, ow, C), np.uint8)
    ys, xs = np.mgrid[0:oh, 0:ow]
    zero = np.zeros((oh, ow), bool)
    if tw > 0:
        xs, ys = xs % tw, ys % th
        if frame:
            zero |= (xs == 0) | (ys == 0) | (xs == tw - 1) | (ys == th - 1)
    if flip:
        xs = ow - 1 - xs
    sx, sy = xs + ox, ys + oy
    if not clamp:
        zero |= (sx < 0) | (sx >= sw) | (sy < 0) | (sy >= sh)
    res = src[np.clip(sy, 0, sh - 1), np.clip(sx, 0, sw - 1)]
    res[zero] = 0
    out[...] = res
    return 0


def tb_img_affine_u8(src, H, W, C, out, matrix6, bicubic, stream):
    from oracle import pil_affine_ref as A
    src, out = _arr(src, (H, W, C), np.uint8), _arr(out, (H, W, C), np.uint8)
    m = tuple(_arr(matrix6, (6,), np.float64))
    out[...] = (A.affine_bicubic if bicubic else A.affine_nearest)(src, m)
    return 0


def tb_img_grayscale_u8(src, out, npix, stream):
    from oracle import pil_affine_ref as A
    src, out = _arr(src, (npix, 1, 3), np.uint8), _arr(out, (npix, 1, 3), np.uint8)
    out[...] = A.grayscale(src)
    return 0


FAKES = {f.__name__: f for f in (tb_resize_crop_normalize_u8, tb_img_gather_u8, tb_img_affine_u8, tb_img_grayscale_u8)}


def install(monkeypatch):
    from textboost_b200 import _cabi, image_ops

    def call(name, *args):
        # what ctypes itself would enforce with the declared argtypes: arity, Python ints for the integer slots (a numpy
        # integer is rejected by c_int), numbers for c_float, pointers / None / addresses for c_void_p
        sig = _cabi._SIGNATURES[name]
        assert len(args) == len(sig), (name, len(args), len(sig))
        for i, (a, t) in enumerate(zip(args, sig)):
            if t in (ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t):
                assert type(a) is int, (name, i, type(a))
            elif t is ctypes.c_float:
                assert type(a) in (float, int), (name, i, type(a))
            elif t is ctypes.c_void_p:
                assert a is None or type(a) is int or isinstance(a, (ctypes.c_void_p, ctypes.Array, ctypes._Pointer)), \
                    (name, i, type(a))
        rc = FAKES[name](*args)
        assert rc == 0

    monkeypatch.setattr(_cabi, "call", call)
    monkeypatch.setattr(_cabi, "stream_ptr", lambda: ctypes.c_void_p(0))
    monkeypatch.setattr(image_ops, "_require_cuda", lambda t, what: None)
