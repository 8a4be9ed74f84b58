"""Byte-exact GPU image tail: tb_resize_crop_normalize_u8 against Pillow + torchvision themselves
(the installed libraries travel with the image) — identical bytes after the resize, identical float bits after the
normalisation — and the training CLI with --gpu_image_transforms.

First seen green on hardware at the end of round 1; the arithmetic and indexing are also pinned on the CPU by
tests/test_resample_cpu.py.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(built_lib):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")


@pytest.mark.parametrize("w,h,size,center", [(90, 70, 32, False), (70, 90, 32, True), (64, 64, 64, False),
                                             (1024, 768, 512, False), (700, 1000, 512, True), (300, 300, 512, False)])
def test_kernel_equals_pillow_and_torchvision_bitwise(w, h, size, center):
    from PIL import Image
    from torchvision.transforms import v2
    from textboost_b200 import image_ops
    a = np.random.RandomState(w + h).randint(0, 256, (h, w, 3), dtype=np.uint8)
    resized = v2.Resize(size, interpolation=v2.InterpolationMode.LANCZOS)(Image.fromarray(a))
    if center:
        top = max(0, int(round((resized.height - size) / 2.0)))
        left = max(0, int(round((resized.width - size) / 2.0)))
    else:
        torch.manual_seed(w)
        top, left, _, _ = v2.RandomCrop.get_params(resized, (size, size))
    window = v2.functional.crop(resized, top, left, size, size)
    want = v2.Compose([v2.ToImage(), v2.ToDtype(torch.float, scale=True), v2.Normalize((0.5,) * 3, (0.5,) * 3)])(window)
    got = image_ops.resize_crop_normalize(torch.from_numpy(a).to(dev), image_ops.shorter_side_size(w, h, size),
                                          top, left, size, size)
    assert got.shape == (3, size, size) and got.dtype == torch.float32
    assert torch.equal(got.cpu(), want)


def test_batch_to_pixel_values_matches_host_dataset(tmp_path):
    import make_augment_golden as G
    from textboost_b200 import augment, dataset, image_ops
    from textboost_b200.synthetic import LiteralTokenizer
    inst, _, _ = G.write_image_dirs(str(tmp_path))
    concepts = [{"instance_data_dir": inst, "instance_token": "<sks> dog"}]

    def batch(device_transforms):
        ds = dataset.TextBoostDataset(concepts, LiteralTokenizer(), template="textboost", size=32,
                                      augment_pipe=augment.PairedAugmentation(**G.PIPES[2]),
                                      device_transforms=device_transforms, cache_decoded=device_transforms)
        G.seed_all(4)
        return dataset.TextBoostDataset.collate_fn([ds[i] for i in range(5)], False)

    host, devb = batch(False), batch(True)
    px = image_ops.batch_to_pixel_values(devb["sources"], dev)
    assert torch.equal(px.cpu(), host["pixel_values"]) and torch.equal(host["input_ids"], devb["input_ids"])


def test_cli_with_gpu_image_transforms(tmp_path):
    import make_augment_golden as G
    import train_textboost as T
    from textboost_b200 import synthetic
    ck = str(tmp_path / "model")
    synthetic.write_pretrained(ck, "tiny", seed=12, vae_channels=(64, 64, 128, 128))
    imgs = tmp_path / "dog"
    imgs.mkdir()
    for i, size in enumerate([(160, 140), (128, 128), (150, 200)]):
        G.make_image(size, i).save(imgs / f"{i}.png")
    jl = tmp_path / "prompts.jsonl"
    with open(jl, "w") as f:
        for i in range(4):
            f.write(json.dumps({"input": f"a thing {i}", "output": "NONE"}) + "\n")

    def run(out, extra):
        return T.main(T.parse_args([
            "--pretrained_model_name_or_path", ck, "--output_dir", str(tmp_path / out), "--instance_data_dir",
            str(imgs), "--resolution", "128", "--train_batch_size", "2", "--max_train_steps", "5", "--learning_rate",
            "1e-3", "--mixed_precision", "fp16", "--augment", "pda", "--augment_inversion", "--template", "textboost",
            "--prior_prompts_file", str(jl), "--log_every", "1", "--seed", "21", "--dataloader_num_workers", "0",
            *extra]))

    loss_gpu = run("a", ["--gpu_image_transforms"])
    loss_host = run("b", [])
    assert loss_gpu == loss_gpu and os.path.exists(tmp_path / "a" / "dog.bin")
    # same seeds, same pixels, same latents noise stream: the two runs agree to the step's run-to-run noise
    assert abs(loss_gpu - loss_host) < 2e-2 * abs(loss_host)


# ------------------------------------------------------------------ deferred augmentation primitives (csrc/augment.cu)
def _plan_of(img):
    from textboost_b200.image_plan import ImagePlan
    return ImagePlan(torch.from_numpy(np.array(img, dtype=np.uint8)))


@pytest.mark.parametrize("size", [(64, 64), (96, 72), (50, 81), (512, 384)])
def test_each_primitive_kernel_is_byte_exact(size):
    """run_plan on single-primitive plans against the pinned oracles (tests/plan_standin.py)."""
    import make_augment_golden as G
    import plan_standin
    from torchvision.transforms.v2.functional._geometry import _get_inverse_affine_matrix
    from textboost_b200.image_plan import run_plan
    base = _plan_of(G.make_image(size, 5))
    w, h = size
    m_scale = _get_inverse_affine_matrix([w * 0.5, h * 0.5], 0.0, [0.0, 0.0], 0.61, [0.0, 0.0])
    m_zoom = _get_inverse_affine_matrix([w * 0.5, h * 0.5], 0.0, [0.0, 0.0], 1.37, [0.0, 0.0])
    m_shift = _get_inverse_affine_matrix([w * 0.5, h * 0.5], 0.0, [-11.0, 0.0], 1.0, [0.0, 0.0])
    plans = [base.pad_edge(7, 0), base.pad_edge(3, 9), base.crop((4, 6, 30, 29)), base.center_crop(20, 24),
             base.center_crop(h + 6, w - 5), base.transpose(Image_FLIP()), base.grayscale(), base.collage(2),
             base.collage(3), base.affine(m_scale, "bicubic"), base.affine(m_zoom, "bicubic"),
             base.affine(m_shift, "nearest"), base.resize((w // 2, h // 3), Image_BICUBIC()),
             base.resize((w + 13, h + 5), Image_BICUBIC())]
    for plan in plans:
        got = run_plan(plan, dev)
        assert got.dtype == torch.uint8 and tuple(got.shape) == (plan.height, plan.width, 3)
        assert np.array_equal(got.cpu().numpy(), plan_standin.run(plan)), plan


def Image_FLIP():
    from PIL import Image
    return Image.FLIP_LEFT_RIGHT


def Image_BICUBIC():
    from PIL import Image
    return Image.BICUBIC


@pytest.mark.parametrize("cfg", [1, 2, 4])
def test_recorded_pipeline_on_gpu_equals_eager_pil(cfg):
    import make_augment_golden as G
    from textboost_b200 import augment
    from textboost_b200.image_plan import run_plan
    eager_pipe, plan_pipe = (augment.PairedAugmentation(**G.PIPES[cfg]) for _ in range(2))
    G.seed_all(5)
    eager = [eager_pipe(G.make_image((128, 128), i), "a <sks> dog") for i in range(16)]
    G.seed_all(5)
    plans = [plan_pipe(_plan_of(G.make_image((128, 128), i)), "a <sks> dog") for i in range(16)]
    for (img, prompt, _), (plan, prompt_p, _) in zip(eager, plans):
        assert prompt == prompt_p
        assert np.array_equal(run_plan(plan, dev).cpu().numpy(), np.asarray(img)), plan


def test_cli_with_gpu_augment(tmp_path):
    import make_augment_golden as G
    import train_textboost as T
    from textboost_b200 import synthetic
    ck = str(tmp_path / "model")
    synthetic.write_pretrained(ck, "tiny", seed=13, vae_channels=(64, 64, 128, 128))
    imgs = tmp_path / "dog"
    imgs.mkdir()
    for i, size in enumerate([(160, 160), (128, 128), (200, 200)]):
        G.make_image(size, i).save(imgs / f"{i}.png")
    jl = tmp_path / "prompts.jsonl"
    with open(jl, "w") as f:
        for i in range(4):
            f.write(json.dumps({"input": f"a thing {i}", "output": "NONE"}) + "\n")

    def run(out, extra):
        return T.main(T.parse_args([
            "--pretrained_model_name_or_path", ck, "--output_dir", str(tmp_path / out), "--instance_data_dir",
            str(imgs), "--resolution", "128", "--train_batch_size", "2", "--max_train_steps", "5", "--learning_rate",
            "1e-3", "--mixed_precision", "fp16", "--augment", "pda", "--augment_inversion", "--augment_p", "0.9",
            "--template", "textboost", "--prior_prompts_file", str(jl), "--log_every", "1", "--seed", "22",
            "--dataloader_num_workers", "0", *extra]))

    loss_gpu = run("a", ["--gpu_augment"])
    loss_host = run("b", [])
    assert loss_gpu == loss_gpu and abs(loss_gpu - loss_host) < 2e-2 * abs(loss_host)
