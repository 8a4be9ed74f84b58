"""Deferred augmentation: the ops of textboost_b200.augment run on an ImagePlan record primitives
instead of calling PIL; executing the record with the pinned oracles (tests/plan_standin.py) must reproduce, byte for
byte, what the same ops — same seeds, same random streams — produce eagerly with PIL / torchvision, and the captions
and the state of the random streams must be identical.  The GPU executor runs the same primitives."""
import random

import numpy as np
import pytest
import torch
from PIL import Image

import make_augment_golden as G
import plan_standin


def _plan_of(img):
    from textboost_b200.image_plan import ImagePlan
    return ImagePlan(torch.from_numpy(np.array(img, dtype=np.uint8)))


DEFERRED_OPS = ["adjust_scale", "horizontal_flip", "horizontal_translate", "grayscale", "crop", "square_photo_collage"]


@pytest.mark.parametrize("name", DEFERRED_OPS)
@pytest.mark.parametrize("size", [(64, 64), (96, 72), (50, 81)])
def test_each_op_recorded_equals_eager(name, size):
    from textboost_b200 import augment
    fn = getattr(augment, name)
    for inversion in (False, True):
        for seed in range(5):
            img = G.make_image(size, seed)
            G.seed_all(seed * 7 + 1)
            eager, prompt_e = fn(img, "a photo of a <sks> dog", inversion)
            after_e = (float(np.random.random()), random.random())
            G.seed_all(seed * 7 + 1)
            plan, prompt_p = fn(_plan_of(img), "a photo of a <sks> dog", inversion)
            after_p = (float(np.random.random()), random.random())
            assert prompt_e == prompt_p and after_e == after_p
            assert plan.size == eager.size
            assert np.array_equal(plan_standin.run(plan), np.asarray(eager)), (name, size, inversion, seed, plan)


def test_ops_outside_the_pipeline_refuse_plans():
    from textboost_b200 import augment
    plan = _plan_of(G.make_image((32, 32), 0))
    for name in ("rotate", "adjust_brightness", "jpeg_compression"):
        with pytest.raises(NotImplementedError):
            getattr(augment, name)(plan, "a dog", False)
    with pytest.raises(NotImplementedError):
        plan.transpose(Image.FLIP_TOP_BOTTOM)
    with pytest.raises(NotImplementedError):
        plan.crop((-1, 0, 10, 10))
    assert plan.resize((32, 32), Image.BICUBIC) is plan and plan.copy() is plan
    assert plan.pad_edge(2, 3).size == (36, 38) and plan.collage(3).size == (96, 96)


@pytest.mark.parametrize("cfg", range(len(G.PIPES)))
def test_pipeline_streams_recorded_equal_eager(cfg):
    """One continuous 24-image stream per PairedAugmentation configuration (the golden configurations): captions,
    bytes and final random-stream state agree between eager PIL execution and record-then-execute."""
    from textboost_b200 import augment
    kw = G.PIPES[cfg]
    eager_pipe, plan_pipe = augment.PairedAugmentation(**kw), augment.PairedAugmentation(**kw)
    G.seed_all(123)
    eager = [eager_pipe(G.make_image(G.SIZES[0], i), "a <sks> dog") for i in range(24)]
    end_e = (float(np.random.random()), random.random())
    G.seed_all(123)
    plans = [plan_pipe(_plan_of(G.make_image(G.SIZES[0], i)), "a <sks> dog") for i in range(24)]
    end_p = (float(np.random.random()), random.random())
    assert end_e == end_p
    n_ops = 0
    for (img, prompt, mask), (plan, prompt_p, mask_p) in zip(eager, plans):
        assert prompt == prompt_p and mask is None and mask_p is None
        assert np.array_equal(plan_standin.run(plan), np.asarray(img)), plan
        n_ops += len(plan.ops)
    assert n_ops > 0 or kw["p"] == 0


def _gather_standin(src, out_h, out_w, ox, oy, clamp, flip, tw, th, frame):
    """numpy transcription of img_gather_u8_kernel."""
    H, W, _ = src.shape
    ys, xs = np.mgrid[0:out_h, 0:out_w]
    zero = np.zeros((out_h, out_w), bool)
    if tw > 0:
        xs, ys = xs % tw, ys % th
        if frame:
            zero |= (xs == 0) | (ys == 0) | (xs == tw - 1) | (ys == th - 1)
    if flip:
        xs = out_w - 1 - xs
    sx, sy = xs + ox, ys + oy
    if clamp:
        sx, sy = np.clip(sx, 0, W - 1), np.clip(sy, 0, H - 1)
    else:
        zero |= (sx < 0) | (sx >= W) | (sy < 0) | (sy >= H)
        sx, sy = np.clip(sx, 0, W - 1), np.clip(sy, 0, H - 1)
    out = src[sy, sx]
    out[zero] = 0
    return out


@pytest.mark.parametrize("size", [(40, 40), (52, 31), (31, 52)])
def test_gather_parameters_of_the_index_primitives(size):
    """image_plan.gather_params + the gather kernel's index rule == the oracle for pad / crop / centre crop (incl. the
    zero-padding case) / mirror / framed collage."""
    from textboost_b200.image_plan import gather_params, op_output_size
    a = np.asarray(G.make_image(size, 3))
    H, W, _ = a.shape
    ops = [("pad_edge", 5, 0), ("pad_edge", 3, 7), ("crop", 4, 6, 20, 17), ("center_crop", 20, 24),
           ("center_crop", W, H), ("center_crop", H + 5, W - 3), ("center_crop", 61, 63), ("flip_lr",), ("collage", 2),
           ("collage", 3)]
    for op in ops:
        w, h = op_output_size(op, W, H)
        got = _gather_standin(a, h, w, *gather_params(op, W, H))
        assert np.array_equal(got, plan_standin.apply_op(a, op)), op
    with pytest.raises(ValueError):
        gather_params(("grayscale",), W, H)


def test_dataset_device_augment_mode_is_the_same_data(tmp_path):
    """TextBoostDataset(device_augment=True): items carry an ImagePlan; executing it with the pinned oracles and
    finishing with the kernel arithmetic gives the host pipeline's "image" bit for bit; prompts, crop positions and
    the random streams are unchanged."""
    from test_resample_cpu import kernel_standin
    from textboost_b200 import augment, dataset
    from textboost_b200.image_plan import ImagePlan
    from textboost_b200.synthetic import LiteralTokenizer
    inst, _, cls = G.write_image_dirs(str(tmp_path))
    concepts = [{"instance_data_dir": inst, "instance_token": "<sks> dog"}]

    def items(device_augment, prior, center=False):
        ds = dataset.TextBoostDataset(concepts, LiteralTokenizer(), template="textboost", size=32,
                                      augment_pipe=augment.PairedAugmentation(**G.PIPES[2]), class_token="dog",
                                      prior_data_root=cls if prior else None, device_augment=device_augment,
                                      center_crop=center, augment_prior=prior and center)
        G.seed_all(31)
        out = [ds[i] for i in range(9)]
        return out, (float(np.random.random()), random.random(), float(torch.rand(1)))

    for prior, center in ((False, False), (True, False), (True, True)):
        host, end_h = items(False, prior, center)
        plan, end_p = items(True, prior, center)
        assert end_h == end_p
        n_ops = 0
        for a, b in zip(host, plan):
            assert isinstance(b["source"], ImagePlan) and "image" not in b
            assert torch.equal(a["input_ids"], b["input_ids"]) and a["crop_top_left"] == b["crop_top_left"]
            assert a["original_size"] == b["original_size"]
            top, left = b["crop_top_left"]
            got, _ = kernel_standin(plan_standin.run(b["source"]), b["resize_to"], top, left, 32, 32)
            assert torch.equal(torch.from_numpy(got), a["image"])
            n_ops += len(b["source"].ops)
            if prior:  # class images are deferred too: Lanczos resize + first crop recorded, tail on top
                assert isinstance(b["class_source"], ImagePlan) and b["class_source"].size == (32, 32)
                assert torch.equal(a["class_input_ids"], b["class_input_ids"])
                assert a["class_crop_top_left"] == b["class_crop_top_left"]
                top, left = b["class_crop_top_left"]
                got, _ = kernel_standin(plan_standin.run(b["class_source"]), b["class_resize_to"], top, left, 32, 32)
                assert torch.equal(torch.from_numpy(got), a["class_image"])
        assert n_ops > 5
    batch = dataset.TextBoostDataset.collate_fn(plan[:3], True)
    assert len(batch["sources"]) == 6 and isinstance(batch["sources"][0]["source"], ImagePlan)


def test_host_plumbing_of_the_image_calls_end_to_end(tmp_path, monkeypatch):
    """batch_to_pixel_values / run_plan / resize_u8 / resize_crop_normalize driven on CPU tensors through a fake
    `_cabi.call` that works on the raw pointers it is given (tests/cabi_standin.py): argument order of the C calls,
    buffer sizes, table upload, the plan executor's op mapping — the batch must equal the host pipeline's bits."""
    import cabi_standin
    from textboost_b200 import augment, dataset, image_ops, image_plan
    from textboost_b200.synthetic import LiteralTokenizer
    cabi_standin.install(monkeypatch)
    inst, _, cls = G.write_image_dirs(str(tmp_path))
    concepts = [{"instance_data_dir": inst, "instance_token": "<sks> dog"}]

    def batch(mode, prior):
        ds = dataset.TextBoostDataset(concepts, LiteralTokenizer(), template="textboost", size=32,
                                      augment_pipe=augment.PairedAugmentation(**G.PIPES[2]), class_token="dog",
                                      prior_data_root=cls if prior else None,
                                      device_transforms=mode == "tail", cache_decoded=mode == "tail",
                                      device_augment=mode == "plan")
        G.seed_all(17)
        return dataset.TextBoostDataset.collate_fn([ds[i] for i in range(6)], prior)

    for prior in (False, True):
        host = batch("host", prior)
        for mode in ("tail", "plan"):
            b = batch(mode, prior)
            px = image_ops.batch_to_pixel_values(b["sources"], "cpu")
            assert px.shape == host["pixel_values"].shape and torch.equal(px, host["pixel_values"]), (mode, prior)
            assert torch.equal(b["input_ids"], host["input_ids"])
    # one decoded base per file stays cached under its path key
    assert all(k[1].endswith(".png") for k in image_plan._device_bases)
    # resize_u8 alone, down and up, against Pillow
    a = np.asarray(G.make_image((60, 44), 2))
    for size in ((30, 22), (75, 50)):
        got = image_ops.resize_u8(torch.from_numpy(a.copy()), size, "bicubic").numpy()
        assert np.array_equal(got, np.asarray(Image.fromarray(a).resize(size, Image.BICUBIC)))


# ---------------------------------------------------------------- the CUDA kernel SOURCE itself, compiled for the host
def test_kernel_source_on_host_end_to_end_equals_the_host_pipeline(tmp_path, monkeypatch):
    """Same as the plumbing test above, but the C calls now run the actual kernel source of csrc/image.cu and
    csrc/augment.cu (host build, one emulated thread: tests/kernel_host_emulation.py): the lines that run on the GPU
    reproduce PIL + torchvision bit for bit through the whole image path, with and without the class-image half."""
    import kernel_host_emulation
    from textboost_b200 import augment, dataset, image_ops
    from textboost_b200.synthetic import LiteralTokenizer
    kernel_host_emulation.install(monkeypatch)
    inst, _, cls = G.write_image_dirs(str(tmp_path))
    concepts = [{"instance_data_dir": inst, "instance_token": "<sks> dog"}]

    def batch(mode, prior, cfg):
        ds = dataset.TextBoostDataset(concepts, LiteralTokenizer(), template="textboost", size=32,
                                      augment_pipe=augment.PairedAugmentation(**G.PIPES[cfg]), class_token="dog",
                                      prior_data_root=cls if prior else None, augment_prior=prior,
                                      device_transforms=mode == "tail", cache_decoded=mode == "tail",
                                      device_augment=mode == "plan")
        G.seed_all(23 + cfg)
        return dataset.TextBoostDataset.collate_fn([ds[i] for i in range(8)], prior)

    n_ops = 0
    for prior, cfg in ((False, 2), (True, 2), (False, 1), (False, 4)):
        host = batch("host", prior, cfg)
        for mode in ("tail", "plan"):
            b = batch(mode, prior, cfg)
            px = image_ops.batch_to_pixel_values(b["sources"], "cpu")
            assert torch.equal(px, host["pixel_values"]), (mode, prior, cfg)
            if mode == "plan":
                n_ops += sum(len(s["source"].ops) for s in b["sources"])
    assert n_ops > 20


@pytest.mark.parametrize("size", [(64, 64), (96, 72), (50, 81), (200, 150)])
def test_kernel_source_on_host_every_primitive_vs_pillow(size, monkeypatch):
    """Every primitive through run_plan with the host build of the kernel source, against the pinned oracles — and the
    affine / resize ones directly against PIL."""
    import kernel_host_emulation
    from PIL import Image
    from torchvision.transforms import v2
    from torchvision.transforms.v2.functional._geometry import _get_inverse_affine_matrix
    from textboost_b200.image_plan import run_plan
    kernel_host_emulation.install(monkeypatch)
    img = G.make_image(size, 5)
    base = _plan_of(img)
    w, h = size
    mats = {s: _get_inverse_affine_matrix([w * 0.5, h * 0.5], 0.0, [0.0, 0.0], s, [0.0, 0.0]) for s in (0.41, 0.9, 1.37)}
    m_shift = _get_inverse_affine_matrix([w * 0.5, h * 0.5], 0.0, [-11.0, 0.0], 1.0, [0.0, 0.0])
    plans = [base.pad_edge(7, 0), base.pad_edge(3, 9), base.crop((4, 6, 30, 29)), base.center_crop(20, 24),
             base.center_crop(h + 6, w - 5), base.transpose(Image.FLIP_LEFT_RIGHT), base.grayscale(), base.collage(2),
             base.collage(3), base.affine(m_shift, "nearest"), base.resize((w // 2, h // 3), Image.BICUBIC),
             base.resize((w + 13, h + 5), Image.BICUBIC), base.resize((w // 2, h // 2), Image.LANCZOS)]
    plans += [base.affine(m, "bicubic") for m in mats.values()]
    for plan in plans:
        got = run_plan(plan, "cpu").numpy()
        assert np.array_equal(got, plan_standin.run(plan)), plan
    for s, m in mats.items():
        ref = v2.functional.affine(img, angle=0, translate=(0, 0), scale=s, shear=0, interpolation=Image.BICUBIC)
        assert np.array_equal(run_plan(base.affine(m, "bicubic"), "cpu").numpy(), np.asarray(ref)), s
    ref = np.asarray(img.resize((w // 2, h // 2), Image.LANCZOS))
    assert np.array_equal(run_plan(base.resize((w // 2, h // 2), Image.LANCZOS), "cpu").numpy(), ref)


def test_cli_image_batches_gpu_augment_equals_host_batches(tmp_path, monkeypatch):
    """train_textboost.build_image_batches with --gpu_augment (no workers) vs the plain host loader with
    --dataloader_num_workers 0: identical input_ids and, through the kernel source on the host, identical pixel_values —
    what tests/test_gpu_zz_image_tail.py::test_cli_with_gpu_augment relies on."""
    import kernel_host_emulation
    import train_textboost as T
    from textboost_b200 import image_ops
    from textboost_b200.synthetic import LiteralTokenizer
    kernel_host_emulation.install(monkeypatch)
    d = tmp_path / "imgs"
    d.mkdir()
    for i, size in enumerate([(80, 64), (64, 64), (70, 90)]):
        G.make_image(size, i).save(d / f"{i}.png")

    def batches(extra):
        args = T.parse_args(["--pretrained_model_name_or_path", "x", "--instance_data_dir", str(d), "--resolution", "32",
                             "--train_batch_size", "2", "--augment", "pda", "--augment_inversion", "--augment_p", "0.9",
                             "--seed", "5", "--template", "textboost", "--dataloader_num_workers", "0", *extra])
        args.concepts_list = [{"instance_token": ["<dog>"], "instance_data_dir": str(d)}]
        G.seed_all(5)
        it = T.build_image_batches(args, LiteralTokenizer(), 0, 1)
        return [next(it) for _ in range(4)], T.RUN_INFO["dataloader_workers"]

    host, w_host = batches([])
    plan, w_plan = batches(["--gpu_augment"])
    assert w_host == 0 and w_plan == 0
    args = T.parse_args(["--pretrained_model_name_or_path", "x", "--gpu_augment"])
    assert args.gpu_image_transforms is True
    for a, b in zip(host, plan):
        assert torch.equal(a["input_ids"], b["input_ids"])
        assert torch.equal(image_ops.batch_to_pixel_values(b["sources"], "cpu"), a["pixel_values"])
