"""Byte-exact image tail: Pillow's antialiased 8-bit resize restated (oracle/pil_resample_ref.py) and
its product-side weight tables / window geometry (textboost_b200/image_ops.py), pinned against the installed Pillow and
torchvision themselves — the third-party code that holds this arithmetic for the reference
(/root/reference/textboost/dataset.py:326-351).  The CUDA kernel consumes exactly these tables with the same integer
arithmetic; `kernel_standin` below is that arithmetic in numpy with the kernel's indexing (row0 / crop window)."""
import numpy as np
import pytest
import torch
from PIL import Image
from torchvision.transforms import v2

CASES = [((64, 48), (32, 24)), ((100, 70), (37, 53)), ((33, 47), (80, 91)), ((128, 128), (64, 64)),
         ((257, 131), (300, 153)), ((300, 200), (96, 64)), ((50, 50), (50, 20)), ((31, 90), (31, 45))]
PIL_FILTER = {"lanczos": Image.LANCZOS, "bicubic": Image.BICUBIC}


def _img(w, h, seed):
    return np.random.RandomState(seed).randint(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("filt", ["lanczos", "bicubic"])
@pytest.mark.parametrize("src,dst", CASES)
def test_oracle_resize_is_byte_exact_vs_pillow(src, dst, filt):
    from oracle import pil_resample_ref as R
    a = _img(*src, seed=src[0] + dst[0])
    ref = np.asarray(Image.fromarray(a).resize(dst, PIL_FILTER[filt]))
    assert np.array_equal(R.resize(a, dst, filt), ref)


def test_oracle_grayscale_ramp_and_extremes():
    """Flat, saturated and single-pixel inputs: clipping of the negative Lanczos lobes, 1-pixel axes."""
    from oracle import pil_resample_ref as R
    for a in (np.zeros((9, 7, 3), np.uint8), np.full((9, 7, 3), 255, np.uint8),
              np.tile(np.array([0, 255], np.uint8).repeat(3).reshape(1, 2, 3), (8, 5, 1)), _img(1, 13, 0), _img(13, 1, 1)):
        for dst in ((5, 4), (20, 17), (1, 1)):
            ref = np.asarray(Image.fromarray(a).resize(dst, Image.LANCZOS))
            assert np.array_equal(R.resize(a, dst, "lanczos"), ref), (a.shape, dst)


def test_product_tables_equal_oracle_tables():
    from oracle import pil_resample_ref as R
    from textboost_b200 import image_ops
    for filt in ("lanczos", "bicubic"):
        for n_in, n_out in [(1024, 512), (768, 512), (683, 512), (512, 512), (100, 37), (33, 80), (7, 3), (3, 7),
                            (1500, 512), (513, 512)]:
            b, k, ks = image_ops.resample_tables(n_in, n_out, filt)
            rb, rk, rks = R.coefficients(n_in, n_out, filt)
            assert ks == rks and np.array_equal(b, rb) and np.array_equal(k, rk), (filt, n_in, n_out)
            assert b.dtype == np.int32 and k.dtype == np.int32
            # weights of a row sum to 2^22 up to the per-tap rounding
            assert np.abs(k.sum(1) - (1 << 22)).max() <= ks
    # an axis whose size does not change is an exact identity pass (Pillow skips it; the kernel may run it)
    b, k, _ = image_ops.resample_tables(64, 64, "lanczos")
    assert all(k[i, i - b[i, 0]] == 1 << 22 and np.abs(k[i]).sum() == 1 << 22 for i in range(64))


@pytest.mark.parametrize("w,h,size", [(90, 70, 32), (70, 90, 32), (64, 64, 32), (1023, 767, 512), (500, 1000, 512),
                                      (33, 32, 32)])
def test_shorter_side_size_matches_torchvision(w, h, size):
    from oracle import pil_resample_ref as R
    from textboost_b200 import image_ops
    out = v2.Resize(size, interpolation=v2.InterpolationMode.LANCZOS)(Image.new("RGB", (w, h)))
    assert image_ops.shorter_side_size(w, h, size) == out.size == R.shorter_side_size(w, h, size)


def kernel_standin(src, out_size, top, left, ch, cw, filt="lanczos"):
    """numpy transcription of tb_resize_crop_normalize_u8 (same tables, same indexing, int32 arithmetic)."""
    from textboost_b200 import image_ops
    H, W, C = src.shape
    bx, kx, _ = image_ops.resample_tables(W, out_size[0], filt)
    by, ky, _ = image_ops.resample_tables(H, out_size[1], filt)
    row0, nrows = image_ops.source_rows(by, top, ch)
    half, bits = np.int32(1 << 21), 22
    mid = np.empty((nrows, cw, C), np.uint8)
    for j in range(cw):
        x0, n = bx[left + j]
        acc = (src[row0:row0 + nrows, x0:x0 + n].astype(np.int32) * kx[left + j, :n, None]).sum(1, dtype=np.int32) + half
        mid[:, j] = np.clip(acc >> bits, 0, 255)
    out = np.empty((C, ch, cw), np.float32)
    u8 = np.empty((ch, cw, C), np.uint8)
    for r in range(ch):
        y0, n = by[top + r]
        acc = (mid[y0 - row0:y0 - row0 + n].astype(np.int32) * ky[top + r, :n, None, None]).sum(0, dtype=np.int32) + half
        v = np.clip(acc >> bits, 0, 255)
        u8[r] = v
        f = v.astype(np.float32) * np.float32(image_ops._SCALE_255)
        out[:, r] = ((f - np.float32(0.5)) / np.float32(0.5)).T
    return out, u8


@pytest.mark.parametrize("w,h,size,center", [(90, 70, 32, False), (70, 90, 32, True), (64, 64, 32, False),
                                             (200, 131, 48, False), (40, 77, 64, True)])
def test_kernel_arithmetic_equals_torchvision_pipeline_bitwise(w, h, size, center):
    """resize(shorter side, LANCZOS) -> crop -> ToImage -> ToDtype(scale) -> Normalize(.5, .5): identical float bits."""
    from textboost_b200 import image_ops
    a = _img(w, h, seed=w * h)
    img = Image.fromarray(a)
    resized = v2.Resize(size, interpolation=v2.InterpolationMode.LANCZOS)(img)
    if center:
        top = max(0, int(round((resized.height - size) / 2.0)))
        left = max(0, int(round((resized.width - size) / 2.0)))
    else:
        torch.manual_seed(w)
        top, left, _, _ = v2.RandomCrop.get_params(resized, (size, size))
    window = v2.functional.crop(resized, top, left, size, size)
    tf = v2.Compose([v2.ToImage(), v2.ToDtype(torch.float, scale=True), v2.Normalize((0.5,) * 3, (0.5,) * 3)])
    want = tf(window)
    got, u8 = kernel_standin(a, image_ops.shorter_side_size(w, h, size), top, left, size, size)
    assert np.array_equal(u8, np.asarray(window))
    assert torch.equal(torch.from_numpy(got), want)


@pytest.mark.parametrize("center_crop,with_aug,prior", [(False, True, False), (True, False, False), (False, True, True)])
def test_dataset_device_transforms_mode_is_the_same_data(tmp_path, center_crop, with_aug, prior):
    """TextBoostDataset(device_transforms=True): same prompts, same crop positions, same random-stream state, and the
    kernel arithmetic applied to its "source" + geometry gives the host mode's "image" bit for bit."""
    import random
    import make_augment_golden as G
    from textboost_b200 import augment, dataset
    from textboost_b200.synthetic import LiteralTokenizer
    inst, inst2, cls = G.write_image_dirs(str(tmp_path))
    concepts = [{"instance_data_dir": inst, "instance_token": "<sks> dog"}]

    def items(device_transforms, cache):
        pipe = augment.PairedAugmentation(**G.PIPES[2]) if with_aug else None
        ds = dataset.TextBoostDataset(concepts, LiteralTokenizer(), template="textboost", size=32,
                                      center_crop=center_crop, augment_pipe=pipe, class_token="dog",
                                      prior_data_root=cls if prior else None, augment_prior=prior and with_aug,
                                      device_transforms=device_transforms, cache_decoded=cache)
        G.seed_all(9)
        out = [ds[i] for i in range(6)]
        return out, (float(np.random.random()), random.random(), float(torch.rand(1))), ds

    host, draws_h, _ = items(False, False)
    devi, draws_d, ds = items(True, True)
    assert draws_h == draws_d
    for a, b in zip(host, devi):
        assert "image" not in b and b["source"].dtype == torch.uint8 and b["source"].shape[2] == 3
        assert a["crop_top_left"] == b["crop_top_left"] and a["original_size"] == b["original_size"]
        assert torch.equal(a["input_ids"], b["input_ids"]) and b["crop_size"] == 32
        top, left = b["crop_top_left"]
        got, _ = kernel_standin(b["source"].numpy(), b["resize_to"], top, left, 32, 32)
        assert torch.equal(torch.from_numpy(got), a["image"])
        if prior:
            top, left = b["class_crop_top_left"]
            got, _ = kernel_standin(b["class_source"].numpy(), b["class_resize_to"], top, left, 32, 32)
            assert torch.equal(torch.from_numpy(got), a["class_image"])
            assert torch.equal(a["class_input_ids"], b["class_input_ids"])
    batch = dataset.TextBoostDataset.collate_fn(devi[:2], prior)
    assert "pixel_values" not in batch and len(batch["sources"]) == (4 if prior else 2)
    assert batch["input_ids"].shape == (4 if prior else 2, 77)
    assert set(batch["sources"][0]) == {"source", "resize_to", "crop_top_left", "crop_size"}
    if prior:
        assert torch.equal(batch["sources"][2]["source"], devi[0]["class_source"])
    assert len(ds._decoded) == 3  # each file decoded once


# ------------------------------------------------------------------ geometry / colour ops of the augmentation
@pytest.mark.parametrize("w,h", [(64, 64), (90, 70), (70, 90), (128, 96), (33, 33)])
def test_affine_and_colour_oracles_are_byte_exact_vs_the_real_ops(w, h):
    """oracle/pil_affine_ref.py against the real augmentation ops (PIL + torchvision underneath) for the same draws:
    adjust_scale (edge pad + bicubic affine + centre crop, incl. the (h, w) swap on non-square images),
    horizontal_translate (edge pad + nearest affine + centre crop), grayscale."""
    import random
    from PIL import ImageOps
    from oracle import pil_affine_ref as A
    from textboost_b200 import augment
    a = _img(w, h, seed=7 * w + h)
    img = Image.fromarray(a)
    assert np.array_equal(A.grayscale(a), np.asarray(ImageOps.grayscale(img).convert("RGB")))
    for seed in range(8):
        np.random.seed(seed)
        random.seed(seed)
        ref, _ = augment.adjust_scale(img, "p", False)
        np.random.seed(seed)
        scale = np.random.uniform(0.34, 1.4)
        assert np.array_equal(A.adjust_scale(a, scale), np.asarray(ref)), ("adjust_scale", seed, scale)
        np.random.seed(seed)
        ref, _ = augment.horizontal_translate(img, "p", False)
        np.random.seed(seed)
        direction = np.random.randint(0, 2)
        shift = int(np.random.uniform(low=0.15, high=0.3) * w)
        got = A.horizontal_translate(a, shift, -1 if direction == 0 else 1)
        assert np.array_equal(got, np.asarray(ref)), ("horizontal_translate", seed, direction, shift)


@pytest.mark.parametrize("scale", [0.34, 0.5, 0.77, 1.0, 1.21, 1.4])
def test_affine_bicubic_oracle_vs_pillow_transform(scale):
    from torchvision.transforms.v2.functional._geometry import _get_inverse_affine_matrix
    from oracle import pil_affine_ref as A
    a = _img(57, 41, seed=int(scale * 100))
    ref = np.asarray(v2.functional.affine(Image.fromarray(a), angle=0, translate=(0, 0), scale=scale, shear=0,
                                          interpolation=Image.BICUBIC))
    m = _get_inverse_affine_matrix([57 * 0.5, 41 * 0.5], 0.0, [0.0, 0.0], scale, [0.0, 0.0])
    assert np.allclose(m, A.scale_matrix(57, 41, scale), rtol=0, atol=1e-12)
    assert np.array_equal(A.affine_bicubic(a, m), ref)
