"""GPU tests of the validation / inference sampler through the C ABI: the fused CFG + DPM-Solver++
step, post_quant_conv and uint8 write-out kernels against their formulas, the VAE decoder engine against
oracle/vae_ref.py, the sampling loop (eager and CUDA-graph replay) against oracle/sampler_ref.py driving the oracle
UNet, and the two CLIs that reach it (train_textboost.py --validation_prompts, inference.py)."""
import os

import pytest
import torch

from conftest import act_dtype, tol_scale

pytestmark = pytest.mark.gpu
F16 = act_dtype()  # fp16; bf16 when the file is re-run under the bf16 policy (tests/test_gpu_bf16_policy.py)
TOLX = tol_scale()  # 1 for fp16, 8 for bf16: rel_l2() reports errors in units of the fp16 bounds written below
dev = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(built_lib):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from textboost_b200 import _cabi
    _cabi.call("tb_check_device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def rel_l2(a, b):
    return (((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()) / TOLX


@pytest.mark.parametrize("v_pred", [False, True])
def test_dpm_cfg_step_kernel(v_pred):
    import ops_standin
    from textboost_b200 import ops
    from textboost_b200.pipeline import DPMSolverMultistepScheduler
    s = DPMSolverMultistepScheduler(prediction_type="v_prediction" if v_pred else "epsilon")
    s.set_timesteps(6)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 4, 16, 16, generator=g) * 10
    xr = x.clone()
    x = x.to(dev)
    m = [torch.zeros_like(x), torch.zeros_like(x)]
    mr = [torch.zeros_like(xr), torch.zeros_like(xr)]
    for i in range(6):
        eps = torch.randn(6, 4, 16, 16, generator=g).to(F16)
        k = s.step_coefficients(i)
        last = i == 5
        uin = None if last else torch.zeros(6, 4, 16, 16, dtype=F16, device=dev)
        uin_r = None if last else torch.zeros(6, 4, 16, 16, dtype=F16)
        args = (7.5, k["alpha_i"], k["sigma_i"], v_pred, k["c_x"], k["c_d0"], k["c_d1"])
        ops.dpm_cfg_step(x, eps.to(dev), m[(i + 1) % 2] if k["c_d1"] else None, m[i % 2], uin, *args)
        ops_standin.dpm_cfg_step(xr, eps, mr[(i + 1) % 2] if k["c_d1"] else None, mr[i % 2], uin_r, *args)
        assert rel_l2(x.cpu(), xr) < 1e-5 and rel_l2(m[i % 2].cpu(), mr[i % 2]) < 1e-5, i
        if not last:
            assert torch.equal(uin[:3], uin[3:]) and rel_l2(uin.cpu(), uin_r) < 1e-3
    assert rel_l2(x.cpu(), mr[5 % 2]) < 1e-5  # the last step returns the data prediction


def test_vae_decode_in_and_image_u8_kernels():
    import ops_standin
    from textboost_b200 import ops
    g = torch.Generator().manual_seed(2)
    lat = torch.randn(3, 4, 6, 10, generator=g)
    w, b = torch.randn(4, 4, generator=g) * 0.5, torch.randn(4, generator=g) * 0.1
    z = ops.vae_decode_in(lat.to(dev), w.to(dev), b.to(dev), 0.18215)
    zr = ops_standin.vae_decode_in(lat, w, b, 0.18215)
    assert z.shape == (3, 4, 6, 10) and z.dtype == F16 and rel_l2(z.cpu(), zr) < 1e-3
    rows = (torch.randn(500, 64, generator=g) * 0.8).to(F16)
    rows[:4, :3] = torch.tensor([[-1.0, 1.0, 0.0], [-3.0, 3.0, 0.5], [1 / 255 - 1, 3 / 255 - 1, 0.25], [0.1, 0.2, 0.3]])
    u8 = ops.image_u8(rows.to(dev), 500, 3)
    ref = ops_standin.image_u8(rows, 500, 3)
    assert u8.dtype == torch.uint8 and u8.shape == (500, 3)
    assert (u8.cpu().int() - ref.int()).abs().max() <= 1  # .5 ties: fp32 v*255 may land a hair either side
    assert (u8.cpu() != ref).float().mean() < 0.01
    assert u8[0].tolist() == [0, 255, 128] or u8[0].tolist() == [0, 255, 127]
    assert u8[1].tolist()[:2] == [0, 255]


def _oracle_vae(seed, kw=None):
    from oracle import vae_ref
    torch.manual_seed(seed)
    ref = vae_ref.AutoencoderKLRef(vae_ref.VAEConfig(**(kw or {})))
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.add_(0.1 * torch.randn_like(p))
            elif n.endswith("bias"):
                p.add_(0.05 * torch.randn_like(p))
    return ref.to(dev).eval().requires_grad_(False)


@pytest.mark.parametrize("B,h,w", [(1, 64, 64), (5, 8, 12)])
def test_vae_decoder_engine_matches_oracle_full_config(B, h, w):
    """SD VAE decoder (49.5 M parameters) at the 512^2 output size, and a ragged-chunk non-square small case.
    Tolerance: fp16 activations vs the fp32 oracle, relative L2 <= 1e-2; uint8 images within 2 levels."""
    from textboost_b200 import vae
    ref = _oracle_vae(21)
    eng = vae.VAEDecoderEngine(vae.VAEConfig(), ref.state_dict())
    lat = torch.randn(B, 4, h, w, generator=torch.Generator().manual_seed(B)).to(dev) * 0.18215 * 2
    with torch.no_grad():
        img_r = ref.decode(lat / 0.18215)
        u8_r = ref.to_uint8(ref.decode_latents(lat))
    img = eng.decode(lat, scaling_factor=0.18215)
    assert img.shape == (B, 3, 8 * h, 8 * w) and torch.isfinite(img).all()
    err = rel_l2(img, img_r)
    print(f"VAE_DECODER_PARITY B={B} {8 * h}x{8 * w} rel_l2 {err:.2e}")
    assert err < 1e-2, err
    u8 = eng.decode_u8(lat)
    d = (u8.int() - u8_r.int()).abs()
    assert u8.shape == (B, 8 * h, 8 * w, 3) and d.max() <= 3 * TOLX and (d > TOLX).float().mean() < 0.01


def _tiny_checkpoint(tmp_path, seed=4):
    from textboost_b200 import synthetic
    ck = str(tmp_path / "model")
    usd, csd = synthetic.write_pretrained(ck, "tiny", seed=seed, vae_channels=(64, 64, 128, 128))
    return ck, usd, csd


def test_sampling_loop_eager_graph_and_oracle(tmp_path):
    """pipeline(...) on a tiny random checkpoint: CUDA-graph replay == eager, and the loop tracks the fp32 oracle
    sampler (oracle UNet, same conditioning) within the fp16-through-8-steps-of-CFG envelope (rel L2 <= 5e-2)."""
    from oracle import harness, sampler_ref, unet_ref
    from textboost_b200 import synthetic
    from textboost_b200.pipeline import DPMSolverMultistepScheduler, StableDiffusionPipeline
    ck, usd, _ = _tiny_checkpoint(tmp_path)
    pipe = StableDiffusionPipeline.from_pretrained(ck, safety_checker=None)
    pipe.scheduler = DPMSolverMultistepScheduler.from_config(pipe.scheduler.config)
    pipe = pipe.to(dev)
    steps, N = 8, 2
    lat0 = torch.randn(N, 4, 16, 16, generator=torch.Generator().manual_seed(9))
    pipe.use_cuda_graph = False
    eager = pipe("a photo of a dog", num_images_per_prompt=N, num_inference_steps=steps, latents=lat0,
                 output_type="latent").images.clone()
    pipe.use_cuda_graph = True
    graphed = pipe("a photo of a dog", num_images_per_prompt=N, num_inference_steps=steps, latents=lat0,
                   output_type="latent").images.clone()
    pipe.use_cuda_graph = False
    eager2 = pipe("a photo of a dog", num_images_per_prompt=N, num_inference_steps=steps, latents=lat0,
                  output_type="latent").images.clone()
    pipe.use_cuda_graph = True
    # run-to-run noise floor: GroupNorm's fp32 atomics, amplified by guidance 7.5 over 8 steps of a random UNet
    floor, gerr = rel_l2(eager2, eager), rel_l2(graphed, eager)
    print(f"SAMPLER_GRAPH eager-vs-eager {floor:.2e} graph-vs-eager {gerr:.2e}")
    assert torch.isfinite(eager).all()
    ucfg, _ = synthetic.model_configs("tiny")
    unet = unet_ref.UNet2DConditionModelRef(harness._unet_cfg(ucfg))
    unet.load_state_dict({k: v.float() for k, v in usd.items()})
    unet = unet.to(dev).eval().requires_grad_(False)
    cond, uncond = pipe.encode_prompt("a photo of a dog", dev, N)
    ref = sampler_ref.sample_latents(lambda x, t, e: unet(x, t, e), cond.float(), uncond.float(), lat0.to(dev),
                                     sampler_ref.DPMSolverMultistepRef(), steps, 7.5)
    err = rel_l2(eager, ref)
    print(f"SAMPLER_PARITY steps={steps} rel_l2 {err:.2e}")
    assert gerr < max(2e-2, 4 * floor), (gerr, floor)
    assert err < 5e-2, err
    # images: PIL, one per requested image, generator-seeded runs reproduce
    g = torch.Generator(device=dev).manual_seed(3)
    imgs = pipe("a photo of a dog", num_images_per_prompt=3, num_inference_steps=4, generator=g).images
    assert len(imgs) == 3 and imgs[0].size == (128, 128) and imgs[0].mode == "RGB"
    with pytest.raises(ValueError):
        pipe("a dog", height=100, width=128)


def test_sampling_loop_with_ddpm_scheduler_vs_oracle(tmp_path):
    """--validation_scheduler DDPMScheduler (train_textboost.py:341-346, 493-495): the ancestral loop of the pipeline
    against the fp32 oracle loop (oracle UNet, oracle/sampler_ref.DDPMRef) with identically seeded CUDA generators --
    both draw one [N,4,h,w] fp32 tensor per step -- within the same 16-bit-through-CFG envelope as the DPM-Solver loop;
    and the CUDA-graph replay of the UNet forward agrees with the eager loop."""
    from oracle import harness, sampler_ref, unet_ref
    from textboost_b200 import synthetic
    from textboost_b200.pipeline import DDPMScheduler, StableDiffusionPipeline
    ck, usd, _ = _tiny_checkpoint(tmp_path)
    pipe = StableDiffusionPipeline.from_pretrained(ck, safety_checker=None)
    pipe.scheduler = DDPMScheduler.from_config(pipe.scheduler.config, variance_type="fixed_small")
    pipe = pipe.to(dev)
    steps, N = 8, 2
    lat0 = torch.randn(N, 4, 16, 16, generator=torch.Generator().manual_seed(9))

    def run(graph):
        pipe.use_cuda_graph = graph
        return pipe("a photo of a dog", num_images_per_prompt=N, num_inference_steps=steps, latents=lat0,
                    generator=torch.Generator(device=dev).manual_seed(21), output_type="latent").images.clone()
    eager, graphed, eager2 = run(False), run(True), run(False)
    floor, gerr = rel_l2(eager2, eager), rel_l2(graphed, eager)
    assert torch.isfinite(eager).all()
    ucfg, _ = synthetic.model_configs("tiny")
    unet = unet_ref.UNet2DConditionModelRef(harness._unet_cfg(ucfg))
    unet.load_state_dict({k: v.float() for k, v in usd.items()})
    unet = unet.to(dev).eval().requires_grad_(False)
    cond, uncond = pipe.encode_prompt("a photo of a dog", dev, N)
    c = pipe.scheduler.config
    ref = sampler_ref.sample_latents(lambda x, t, e: unet(x, t, e), cond.float(), uncond.float(), lat0.to(dev),
                                     sampler_ref.DDPMRef(timestep_spacing=c.timestep_spacing,
                                                         steps_offset=c.steps_offset), steps, 7.5,
                                     generator=torch.Generator(device=dev).manual_seed(21))
    err = rel_l2(eager, ref)
    print(f"DDPM_SAMPLER_PARITY steps={steps} rel_l2 {err:.2e} graph-vs-eager {gerr:.2e} floor {floor:.2e}")
    assert gerr < max(2e-2, 4 * floor), (gerr, floor)
    assert err < 5e-2, err
    imgs = pipe("a photo of a dog", num_images_per_prompt=2, num_inference_steps=4,
                generator=[torch.Generator(device=dev).manual_seed(s) for s in (1, 2)]).images
    assert len(imgs) == 2 and imgs[0].size == (128, 128)


def test_training_cli_validation_and_inference_cli(tmp_path, monkeypatch):
    """train_textboost.py --validation_prompts writes validation_<step>.jpg (train_textboost.py:1213-1228); then
    inference.py loads the adapter + learned embeddings it saved and samples a grid (inference.py:46-113)."""
    from PIL import Image
    import inference as I
    import train_textboost as T
    ck, _, _ = _tiny_checkpoint(tmp_path, seed=6)
    out = str(tmp_path / "out")
    T.main(T.parse_args([
        "--pretrained_model_name_or_path", ck, "--output_dir", out, "--synthetic_data", "--resolution", "128",
        "--train_batch_size", "2", "--max_train_steps", "4", "--learning_rate", "1e-3", "--mixed_precision", "fp16",
        "--augment_inversion", "--validation_prompts", "a <0> in the snow", "photo of <0>", "--validation_steps", "4",
        "--num_validation_images", "2", "--log_every", "2"]))
    assert T.RUN_INFO["validation_steps"][-1] == 4
    grid = Image.open(os.path.join(out, "validation_4.jpg"))
    assert grid.size == (2 * 128, 2 * 128)
    assert {"text_encoder", "dog.bin", "hflip.bin"} <= set(os.listdir(out))
    # --validation_scheduler DDPMScheduler (train_textboost.py:341-346): the ancestral sampler behind the same flag
    out_d = str(tmp_path / "out_ddpm")
    T.main(T.parse_args([
        "--pretrained_model_name_or_path", ck, "--output_dir", out_d, "--synthetic_data", "--resolution", "128",
        "--train_batch_size", "2", "--max_train_steps", "2", "--learning_rate", "1e-3", "--mixed_precision", "fp16",
        "--validation_prompts", "a <0> in the snow", "--validation_steps", "2", "--num_validation_images", "2",
        "--validation_scheduler", "DDPMScheduler", "--seed", "3"]))
    assert Image.open(os.path.join(out_d, "validation_2.jpg")).size == (2 * 128, 128)
    monkeypatch.chdir(tmp_path)
    sheet = str(tmp_path / "sheet.jpg")
    I.main(I.parse_args([out + "/", "--model", ck, "--prompt", "photo of a <dog> dog", "--seeds", "0", "1", "2",
                         "--output", sheet, "--num_inference_steps", "4"]))
    assert Image.open(sheet).size == (3 * 128, 128)
    files = I.main(I.parse_args([out, "--model", ck, "--prompt", "a <dog>", "--seeds", "5",
                                 "--num_inference_steps", "3"]))
    assert files == ["a_<dog>_5.jpg"] and os.path.exists(tmp_path / "a_<dog>_5.jpg")
    with pytest.raises(OSError):
        I.main(I.parse_args([out, "--model", "sd21base"]))
