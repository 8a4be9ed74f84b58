"""The SIMT half of the C ABI on the CPU: the product source of csrc/{norm,elementwise,clip,optim,vae,sampler,image,
augment}.cu — kernels AND extern "C" entry points — built for the host behind a block emulator (every CUDA thread an OS
thread, real barriers / shuffles / atomics; tests/kernel_host_emulation.py) and bound to the product's own ctypes
layer.  `ops.*`, `FusedAdamW` and `C.call` then run unchanged on CPU tensors, so the CPU suite exercises the argument
checks, launch geometry, dispatch and kernel code of 44 entry points against the oracle formulas.  The tcgen05 entry
points (GEMM, conv3x3, attention) need the hardware and are covered by the `-m gpu` tests only."""
import ctypes
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

import kernel_host_emulation as K

# a stuck barrier in the block emulator must fail the test, not hang the suite
pytestmark = pytest.mark.timeout(600, method="thread")


@pytest.fixture()
def abi(monkeypatch):
    return K.install_abi(monkeypatch)


def _h(*shape, seed=0, scale=1.0, shift=0.0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale + shift).half()


def _close(a, b, tol):
    d = (a.float() - b.float()).abs().max().item()
    assert d <= tol * max(1.0, b.float().abs().max().item()), d


@pytest.mark.parametrize("B,HW,C,G,silu", [(2, 40, 64, 32, True), (1, 72, 320, 32, False), (2, 9, 640, 32, True),
                                          # the single-launch group-owner kernel: one group per CTA (cpg 40), two
                                          # groups per CTA (cpg 20 / 60: a vector straddles the two groups)
                                          (2, 12, 1280, 32, True), (3, 10, 640, 32, True), (3, 7, 1920, 32, False)])
def test_groupnorm_entry_points(abi, B, HW, C, G, silu):
    """tb_groupnorm_{fwd,bwd}_f16 incl. the two-group fast path (cpg >= 8) and the generic one (cpg = 2)."""
    from textboost_b200 import ops
    x = _h(B, HW, C, seed=C, scale=1.5, shift=0.3)
    gamma, beta = _h(C, seed=1, scale=0.1, shift=1.0), _h(C, seed=2, scale=0.1)
    y, stats = ops.groupnorm(x, gamma, beta, G, 1e-5, silu)
    xf = x.float().transpose(1, 2).requires_grad_(True)
    ref = F.group_norm(xf, G, gamma.float(), beta.float(), 1e-5)
    ref = F.silu(ref) if silu else ref
    _close(y, ref.detach().transpose(1, 2), 2e-3)
    dy, add = _h(B, HW, C, seed=3), _h(B, HW, C, seed=4)
    dx = ops.groupnorm_bwd(dy, x, gamma, beta, stats, G, 1e-5, silu, add=add)
    ref.backward(dy.float().transpose(1, 2))
    _close(dx, xf.grad.transpose(1, 2) + add.float(), 3e-3)
    # the saved statistics are (sum x, sum x^2) per (image, group) whichever kernel wrote them
    xs = x.float().view(B, HW, G, C // G)
    _close(stats[..., 0], xs.sum((1, 3)), 1e-3)
    _close(stats[..., 1], (xs * xs).sum((1, 3)), 1e-3)


@pytest.mark.parametrize("M,D,R,RPAD", [(11, 768, 12, 16), (5, 1024, 40, 48), (9, 128, 12, 16)])
def test_text_encoder_layernorm_with_lora_glue(abi, M, D, R, RPAD):
    """tb_layernorm_lora_fwd / tb_layernorm_bwd_clip (LayerNorm + LoRA down-projection; LayerNorm backward + the
    down-projection's input-gradient + the fp16 copy) run from the product source on the host."""
    from textboost_b200 import ops
    g = torch.Generator().manual_seed(M + D)
    x = torch.randn(M, D, generator=g) * 2 + 0.3
    gamma, beta = torch.randn(D, generator=g), torch.randn(D, generator=g)
    A = torch.randn(R, D, generator=g) / R
    y_ext = torch.full((M, D + RPAD + 8), 5.0).half()
    stats = ops.layernorm_lora_fwd(x, gamma, beta, A, y_ext, RPAD)
    _close(y_ext[:, :D], F.layer_norm(x, (D,), gamma, beta), 1e-3)
    _close(y_ext[:, D:D + R], y_ext[:, :D].float() @ A.t(), 2e-3)
    assert bool((y_ext[:, D + R:D + RPAD] == 0).all()) and bool((y_ext[:, D + RPAD:] == 5.0).all())
    dy_ext = torch.randn(M, D + RPAD, generator=g).half()
    add = torch.randn(M, D, generator=g)
    xr = x.clone().requires_grad_(True)
    F.layer_norm(xr, (D,), gamma, beta).backward(dy_ext[:, :D].float() + dy_ext[:, D:D + R].float() @ A)
    out16 = torch.empty(M, D).half()
    buf = add.clone()
    dx = ops.layernorm_bwd_clip(dy_ext, x, gamma, stats, add=buf, out=buf, out16=out16, lora_a=A)
    torch.testing.assert_close(dx, xr.grad + add, rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(out16.float(), dx, rtol=1e-3, atol=1e-3)
    dy32 = torch.randn(M, D, generator=g)
    xr2 = x.clone().requires_grad_(True)
    F.layer_norm(xr2, (D,), gamma, beta).backward(dy32)
    torch.testing.assert_close(ops.layernorm_bwd_clip(dy32, x, gamma, stats), xr2.grad, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("M,C,f32", [(13, 320, False), (9, 640, False), (10, 128, False), (11, 768, True), (6, 1024, True)])
def test_layernorm_entry_points(abi, M, C, f32):
    """tb_layernorm_{fwd,bwd}: every lanes-per-row / vectors-per-lane instantiation, fp16 (UNet) and fp32 (CLIP) sets."""
    from textboost_b200 import ops
    dt = torch.float32 if f32 else torch.float16
    g = torch.Generator().manual_seed(C)
    x = (torch.randn(M, C, generator=g) * 1.3 + 0.2).to(dt)
    gamma, beta = (1 + 0.1 * torch.randn(C, generator=g)).to(dt), (0.1 * torch.randn(C, generator=g)).to(dt)
    y, stats = ops.layernorm(x, gamma, beta, eps=1e-5)
    xf = x.float().requires_grad_(True)
    ref = F.layer_norm(xf, (C,), gamma.float(), beta.float(), 1e-5)
    _close(y, ref.detach(), 2e-3)
    dy = torch.randn(M, C, generator=g).half()
    add = torch.randn(M, C, generator=g).to(dt)
    dx = ops.layernorm_bwd(dy, x, gamma, stats, add=add)
    ref.backward(dy.float())
    assert dx.dtype == dt
    _close(dx, xf.grad + add.float(), 3e-3)


def test_mse_and_kpl_entry_points(abi):
    from textboost_b200 import _cabi as C
    from textboost_b200 import ops
    g = torch.Generator().manual_seed(0)
    pred, target = _h(2, 4, 8, 8, seed=1), torch.randn(2, 4, 8, 8, generator=g)
    loss, scale = torch.zeros(1), torch.tensor([128.0])
    dpred = ops.mse_fwd_bwd(pred, target, loss, 0.5, scale)
    pf = pred.float().requires_grad_(True)
    ref = 0.5 * F.mse_loss(pf, target)
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item()) + 1e-7
    _close(dpred, pf.grad * 128.0, 2e-3)
    M, D = 10, 96
    h = torch.randn(M, D, generator=g)
    h0 = h + 0.3 * torch.randn(M, D, generator=g)
    for kind, fn in ((0, lambda a: (1 - F.cosine_similarity(a, h0, dim=-1)).mean()), (1, lambda a: F.mse_loss(a, h0))):
        acc, dh = torch.zeros(1), torch.zeros(M, D)
        C.call("tb_kpl_fwd_bwd", C.ptr(h), C.ptr(h0), M, D, kind, 0.1, C.ptr(scale), C.ptr(acc), C.ptr(dh),
               C.stream_ptr())
        hf = h.clone().requires_grad_(True)
        ref = 0.1 * fn(hf)
        ref.backward()
        assert abs(acc.item() - ref.item()) < 1e-5 * abs(ref.item()) + 1e-8
        torch.testing.assert_close(dh, hf.grad * 128.0, rtol=1e-4, atol=1e-6)


def test_fused_adamw_class_over_the_real_entry_point(abi):
    """textboost_b200.optim.FusedAdamW (product class) -> tb_optim_mix_mask + tb_adamw_fused_step (all three kernels
    incl. the cooperative finishing one) vs GradScaler semantics + clip_grad_norm_(LoRA) + torch.optim.AdamW + the
    added-row renorm of train_textboost.py:1138-1149; then a skipped (inf) step halves the loss scale."""
    from textboost_b200.optim import FusedAdamW
    g = torch.Generator().manual_seed(0)
    D, r, T, n_rows, layers = 16, 4, 3, 2, 2
    n_a, n_b = layers * T * r * D, layers * T * D * r
    n_lora = n_a + n_b
    n = n_lora + n_rows * D
    params = torch.randn(n, generator=g) * 0.5
    params[n_lora:] *= 4  # long rows: the renorm must act
    state = SimpleNamespace(params=params, grads=torch.zeros(n), n_lora=n_lora, n_rows=n_rows, D=D, r=r, n_b=n_b,
                            rows=lambda buf=None: (params if buf is None else buf)[n_lora:].view(n_rows, D),
                            b_segment=lambda buf: buf[n_a:n_lora])
    engine = SimpleNamespace(state=state, tok_base=torch.randn(50, D, generator=g), decay=None)
    opt = FusedAdamW(engine, lr=5e-3, emb_lr=1e-2, mean_norm=2.0, mixing="object")
    ref_lora = torch.nn.Parameter(params[:n_lora].clone())
    ref_rows = torch.nn.Parameter(params[n_lora:].clone())
    ref = torch.optim.AdamW([{"params": [ref_rows], "lr": 1e-2}, {"params": [ref_lora], "lr": 5e-3}], weight_decay=1e-2)
    for step in range(3):
        grad = torch.randn(n, generator=g) * (2.0 if step == 0 else 0.1)
        state.grads.copy_(grad * 65536.0)
        opt.step()
        gl = grad[:n_lora].clone()
        gl[n_a:].view(-1, D, r)[:, 1::2, :] = 0  # --mixing object: odd rows of every lora_B block
        ref_lora.grad, ref_rows.grad = gl, grad[n_lora:].clone()
        norm = torch.nn.utils.clip_grad_norm_([ref_lora], 1.0)
        ref.step()
        with torch.no_grad():
            rows = ref_rows.view(n_rows, D)
            v = rows.norm(dim=-1, keepdim=True)
            rows.copy_(torch.minimum(torch.full_like(v, 2.0), v) / v * rows)
        assert state.grads.abs().max() == 0
        assert abs(opt.state[7].item() - norm.item()) < 1e-4 * norm.item()
        assert abs(opt.added_norm.item() - v.mean().item()) < 1e-4 * v.mean().item()
        torch.testing.assert_close(params[:n_lora], ref_lora.detach(), rtol=2e-5, atol=2e-6)
        torch.testing.assert_close(params[n_lora:], ref_rows.detach(), rtol=2e-5, atol=2e-6)
    assert opt.state[4] == 3 and opt.state[1] == 3 and opt.state[0] == 65536.0
    assert abs(opt.state[5].item() - (1 - 1e-2 * 1e-2) ** 3) < 1e-6  # lazy decay scalar of the frozen rows
    before = params.clone()
    state.grads.copy_(torch.randn(n, generator=g))
    state.grads[7] = float("nan")
    opt.step()
    assert torch.equal(params, before) and opt.state[0] == 32768.0 and opt.state[8] == 1 and opt.state[4] == 3


def test_softmax_and_direct_conv_entry_points(abi):
    from textboost_b200 import ops
    buf = _h(20, 80, seed=5, scale=4.0)
    x = buf[:, :72]
    ref = torch.softmax(x.float(), -1)
    keep = buf[:, 72:].clone()
    ops.softmax_rows_(x)
    _close(x, ref, 1e-3)
    assert torch.equal(buf[:, 72:], keep)
    h, w, b = _h(1, 6, 5, 64, seed=6), _h(4, 64, 3, 3, seed=7, scale=0.1), _h(4, seed=8, scale=0.1)
    y = ops.conv_out(h, w, b)  # one warp per pixel
    _close(y, F.conv2d(h.float().permute(0, 3, 1, 2), w.float(), b.float(), padding=1), 3e-3)
    dy = _h(1, 4, 6, 5, seed=9)
    _close(ops.conv_out_bwd(dy, w), F.conv_transpose2d(dy.float(), w.float(), padding=1).permute(0, 2, 3, 1), 3e-3)


def test_legacy_causal_attention_entry_points(abi):
    """tb_clip_attn_{fwd,bwd}: causal softmax(QK^T/8)V over 77 tokens, head_dim 64 (the CUDA-core kernel kept in the
    ABI), against torch's SDPA and autograd."""
    from textboost_b200 import _cabi as C
    B, L, heads = 1, 21, 2
    D = heads * 64
    qkv = _h(B * L, 3 * D, seed=10, scale=0.5)
    out = torch.empty(B * L, D, dtype=torch.float16)
    C.call("tb_clip_attn_fwd", C.ptr(qkv), C.ptr(out), B, L, D, heads, C.stream_ptr())
    q, k, v = (t.float().view(B, L, heads, 64).transpose(1, 2).requires_grad_(True) for t in qkv.split(D, dim=1))
    ref = F.scaled_dot_product_attention(q, k, v, is_causal=True)
    _close(out.view(B, L, heads, 64).transpose(1, 2), ref.detach(), 3e-3)
    dO = _h(B * L, D, seed=11)
    dqkv = torch.empty(B * L, 3 * D, dtype=torch.float16)
    C.call("tb_clip_attn_bwd", C.ptr(qkv), C.ptr(dO), C.ptr(dqkv), B, L, D, heads, C.stream_ptr())
    ref.backward(dO.float().view(B, L, heads, 64).transpose(1, 2))
    want = torch.cat([t.grad.transpose(1, 2).reshape(B * L, D) for t in (q, k, v)], dim=1)
    _close(dqkv, want, 4e-3)


def test_entry_point_argument_checks_and_error_text(abi):
    """Negative TB_E_* codes with a message through tb_last_error, as the Python layer surfaces them."""
    from textboost_b200 import _cabi as C
    from textboost_b200 import ops
    with pytest.raises(RuntimeError, match="groupnorm: C=60"):
        ops.groupnorm(_h(1, 4, 60), _h(60), _h(60), 32, 1e-5, True)
    with pytest.raises(RuntimeError, match="cols=70"):
        ops.softmax_rows_(_h(4, 70))
    with pytest.raises(RuntimeError, match="second-order update needs m_prev"):
        x = torch.zeros(8)
        ops.dpm_cfg_step(x, torch.zeros(16, dtype=torch.float16), None, torch.zeros(8), None, 7.5, 1.0, 1.0, False, 1.0,
                         1.0, 0.5)
    with pytest.raises(RuntimeError, match="tb_conv_out_f16"):
        ops.conv_out(_h(1, 4, 4, 64), _h(8, 64, 3, 3), _h(8))
    assert C.launch_count > 0


def test_unet_data_movement_and_elementwise_wrappers(abi):
    """ops.* of the UNet's HBM-bound work through their entry points: GEGLU (+bwd), nearest upsample (+bwd), the skip
    concat / split built on tb_copy2d_f16, fp32->fp16 cast into a strided view, Downsample2D lowering (im2col) and its
    backward (zero stuffing), timestep embedding, SiLU, DDPM add_noise / velocity target, conv_in."""
    import ops_standin
    from oracle import ddpm_ref, unet_ref
    from textboost_b200 import ops
    h = _h(6, 48, seed=1)
    hf = h.float().requires_grad_(True)
    ref = hf[:, :24] * F.gelu(hf[:, 24:])
    _close(ops.geglu(h), ref.detach(), 2e-3)
    dg = _h(6, 24, seed=2)
    ref.backward(dg.float())
    _close(ops.geglu_bwd(dg, h), hf.grad, 2e-3)
    x = _h(2, 3, 5, 16, seed=3)
    assert torch.equal(ops.upsample2x(x), ops_standin.upsample2x(x))
    dy = _h(2, 6, 10, 16, seed=4)
    _close(ops.upsample2x_bwd(dy), dy.float().view(2, 3, 2, 5, 2, 16).sum((2, 4)), 2e-3)
    a, b = _h(2, 3, 5, 16, seed=5), _h(2, 3, 5, 24, seed=6)
    cat = ops.concat_channels(a, b)
    assert torch.equal(cat, torch.cat([a, b], -1))
    a2, b2 = ops.split_channels(cat, 16)
    assert torch.equal(a2, a) and torch.equal(b2, b)
    acc = a.reshape(-1, 16).clone()
    ops.copy2d(acc, a.reshape(-1, 16), accumulate=True)
    _close(acc, 2 * a.reshape(-1, 16).float(), 2e-3)
    f = torch.randn(5, 16, generator=torch.Generator().manual_seed(7))
    out = torch.zeros(5, 40, dtype=torch.float16)
    ops.cast_f32_f16(f, out=out[:, 8:24], scale=0.5)
    assert torch.equal(out[:, 8:24], (f * 0.5).half()) and out[:, :8].abs().max() == 0 and out[:, 24:].abs().max() == 0
    xd = _h(2, 8, 6, 16, seed=8)
    assert torch.equal(ops.im2col3x3s2(xd), ops_standin.im2col3x3s2_pad(xd, 1))
    assert torch.equal(ops.im2col3x3s2_pad(xd, 0), ops_standin.im2col3x3s2_pad(xd, 0))
    dz = _h(2, 4, 3, 8, seed=9)
    want = torch.zeros(2, 8, 6, 8, dtype=torch.float16)
    want[:, ::2, ::2] = dz
    assert torch.equal(ops.zero_stuff2x(dz), want)
    t = torch.tensor([0, 1, 500, 999], dtype=torch.int64)
    _close(ops.timestep_embedding(t, 64), unet_ref.timestep_embedding(t, 64), 2e-3)
    s = _h(4, 64, seed=10, scale=3.0)
    _close(ops.silu(s), F.silu(s.float()), 2e-3)
    g = torch.Generator().manual_seed(11)
    x0, eps = torch.randn(4, 4, 8, 8, generator=g), torch.randn(4, 4, 8, 8, generator=g)
    acp = ddpm_ref.alphas_cumprod()
    noisy, target = ops.add_noise(x0, eps, t, acp, v_prediction=True)
    _close(noisy, ddpm_ref.add_noise(x0, eps, t), 2e-3)
    torch.testing.assert_close(target, ddpm_ref.get_velocity(x0, eps, t), rtol=1e-6, atol=1e-6)
    xi, w, bias = _h(2, 4, 6, 5, seed=12), _h(16, 4, 3, 3, seed=13, scale=0.2), _h(16, seed=14, scale=0.1)
    _close(ops.conv_in(xi, w, bias), F.conv2d(xi.float(), w.float(), bias.float(), padding=1).permute(0, 2, 3, 1), 2e-3)


def test_vae_sampler_and_image_wrappers(abi, monkeypatch):
    """The late-built entry points through ops / image_ops / image_plan on CPU tensors: Gaussian sample, post_quant_conv
    input, uint8 write-out, the fused CFG + DPM-Solver++ step, the byte-exact resize / crop / normalise tail and a
    recorded augmentation plan — the last two against PIL / torchvision bit for bit."""
    import make_augment_golden as G
    import numpy as np
    import ops_standin
    import plan_standin
    from PIL import Image
    from torchvision.transforms import v2
    from textboost_b200 import augment, image_ops, ops
    from textboost_b200.image_plan import ImagePlan, run_plan
    monkeypatch.setattr(image_ops, "_require_cuda", lambda t, what: None)
    g = torch.Generator().manual_seed(1)
    rows = torch.randn(3 * 40, 64, generator=g).half()
    eps = torch.randn(3, 4, 40, generator=g)
    lat, mean, std = ops.vae_sample(rows, 3, 40, 4, eps=eps, scaling_factor=0.18215, want_moments=True)
    lat_r, mean_r, std_r = ops_standin.vae_sample(rows, 3, 40, 4, eps=eps, scaling_factor=0.18215, want_moments=True)
    assert torch.equal(mean, mean_r)
    torch.testing.assert_close(std, std_r, rtol=2e-7, atol=0)
    torch.testing.assert_close(lat, lat_r, rtol=1e-6, atol=1e-7)
    latents = torch.randn(2, 4, 5, 7, generator=g)
    w, b = torch.randn(4, 4, generator=g) * 0.5, torch.randn(4, generator=g) * 0.1
    _close(ops.vae_decode_in(latents, w, b, 0.18215), ops_standin.vae_decode_in(latents, w, b, 0.18215), 2e-3)
    px = (torch.randn(300, 64, generator=g) * 0.8).half()
    assert torch.equal(ops.image_u8(px, 300, 3), ops_standin.image_u8(px, 300, 3))
    x = torch.randn(2, 4, 8, 8, generator=g) * 10
    xr = x.clone()
    e = torch.randn(4, 4, 8, 8, generator=g).half()
    m0, m0r = torch.zeros_like(x), torch.zeros_like(x)
    uin, uin_r = torch.zeros(4, 4, 8, 8, dtype=torch.float16), torch.zeros(4, 4, 8, 8, dtype=torch.float16)
    ops.dpm_cfg_step(x, e, None, m0, uin, 7.5, 0.07, 0.99, False, 0.8, 0.3, 0.0)
    ops_standin.dpm_cfg_step(xr, e, None, m0r, uin_r, 7.5, 0.07, 0.99, False, 0.8, 0.3, 0.0)
    torch.testing.assert_close(x, xr, rtol=1e-5, atol=1e-4)
    img = G.make_image((90, 70), 3)
    a = np.asarray(img)
    resized = v2.Resize(32, interpolation=v2.InterpolationMode.LANCZOS)(img)
    window = v2.functional.crop(resized, 0, 5, 32, 32)
    want = v2.Compose([v2.ToImage(), v2.ToDtype(torch.float, scale=True), v2.Normalize((0.5,) * 3, (0.5,) * 3)])(window)
    got = image_ops.resize_crop_normalize(torch.from_numpy(a.copy()), image_ops.shorter_side_size(90, 70, 32), 0, 5, 32, 32)
    assert torch.equal(got, want)
    pipe = augment.PairedAugmentation(**G.PIPES[2])
    G.seed_all(3)
    eager = [pipe(G.make_image((64, 64), i), "a dog")[0] for i in range(6)]
    G.seed_all(3)
    plans = [pipe(ImagePlan(torch.from_numpy(np.array(G.make_image((64, 64), i), dtype=np.uint8))), "a dog")[0]
             for i in range(6)]
    assert sum(len(p.ops) for p in plans) > 3
    for im, plan in zip(eager, plans):
        assert np.array_equal(run_plan(plan, "cpu").numpy(), np.asarray(im)), plan


@pytest.mark.parametrize("nblk,tmask,r", [(3, 0b111, 4), (3, 0b101, 16), (1, 0b1, 8)])
def test_lora_entry_points(abi, nblk, tmask, r):
    """a4: tb_lora_down (xa = LN(x) A^T into the K-extension columns), tb_lora_grad (dB, dA accumulated in fp32 from
    the fused projection gradient) and tb_lora_dx (the A^T path of the input gradient): the reference's three targets
    at rank 4, a subset of the fused q|k|v blocks at rank 16, and the single out_proj block at rank 8."""
    from textboost_b200 import _cabi as C
    g = torch.Generator().manual_seed(0)
    blocks = [b for b in range(nblk) if (tmask >> b) & 1]
    T = len(blocks)
    M, D = 70, 64
    R = T * r
    RPAD = (R + 15) // 16 * 16
    ld = D + RPAD
    A = torch.randn(R, D, generator=g) * 0.3
    y_ext = torch.full((M, ld), 9.0, dtype=torch.float16)
    y_ext[:, :D] = _h(M, D, seed=1)
    C.call("tb_lora_down", C.ptr(y_ext), ld, C.ptr(A), M, D, R, RPAD, C.stream_ptr())
    y = y_ext[:, :D].float()
    _close(y_ext[:, D:D + R], y @ A.t(), 2e-3)
    assert RPAD == R or y_ext[:, D + R:].abs().max() == 0
    dY = _h(M, nblk * D, seed=2)
    dA_ext = torch.zeros(M, ld, dtype=torch.float16)
    dA_ext[:, :D] = _h(M, D, seed=3)
    dA_ext[:, D:D + R] = _h(M, R, seed=4)
    dB, dA = torch.ones(T * D, r), torch.ones(R, D)  # accumulate on top of what is there
    C.call("tb_lora_grad", C.ptr(dY), C.ptr(y_ext), C.ptr(dA_ext), ld, C.ptr(dB), C.ptr(dA), M, nblk, tmask, D, r, 0.5,
           C.stream_ptr())
    xa, dxa = y_ext[:, D:D + R].float(), dA_ext[:, D:D + R].float()
    want_b = torch.cat([0.5 * dY[:, b * D:(b + 1) * D].float().t() @ xa[:, t * r:(t + 1) * r]
                        for t, b in enumerate(blocks)]) + 1
    torch.testing.assert_close(dB, want_b, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dA, dxa.t() @ y + 1, rtol=1e-4, atol=1e-4)
    before = dA_ext[:, :D].float().clone()
    C.call("tb_lora_dx", C.ptr(dA_ext), ld, C.ptr(A), M, D, R, C.stream_ptr())
    _close(dA_ext[:, :D], before + dxa @ A, 2e-3)
    with pytest.raises(RuntimeError, match="tb_lora_grad"):
        C.call("tb_lora_grad", C.ptr(dY), C.ptr(y_ext), C.ptr(dA_ext), ld, C.ptr(dB), C.ptr(dA), M, nblk, tmask, D, 17,
               0.5, C.stream_ptr())


@pytest.mark.parametrize("r", [4, 5])
def test_unet_cross_kv_lora_entry_points(abi, r):
    """tb_unet_lora_fwd / tb_unet_lora_bwd (peft LoRA on attn2.to_k / to_v of all blocks side by side,
    train_textboost.py:712-721) against autograd on the same fp32 adapters; then the third parameter group
    (optim.FlatAdamW: the LoRA learning rate, weight decay, no clipping, :838-841) against torch.optim.AdamW."""
    from textboost_b200 import ops
    from textboost_b200.optim import FlatAdamW
    g = torch.Generator().manual_seed(3 + r)
    M, ctx, widths = 9, 40, (8, 24, 16)
    KV, n_ad = 2 * sum(widths), 2 * len(widths)
    blk, off = [], [0]
    for i, w in enumerate(widths):
        blk += [2 * i] * w + [2 * i + 1] * w
        off += [off[-1] + w, off[-1] + 2 * w]
    blk_t, off_t = torch.tensor(blk, dtype=torch.int32), torch.tensor(off, dtype=torch.int32)
    ehs, kv0, dkv = _h(M, ctx, seed=1), _h(M, KV, seed=2), _h(M, KV, seed=3)
    params = torch.cat([(torch.randn(n_ad * r * ctx, generator=g) / r), 0.05 * torch.randn(KV * r, generator=g)])
    A = params[:n_ad * r * ctx].view(n_ad * r, ctx).clone().requires_grad_(True)
    Bm = params[n_ad * r * ctx:].view(KV, r).clone().requires_grad_(True)
    e32 = ehs.float().requires_grad_(True)
    ref = kv0.float() + torch.cat([1.5 * (e32 @ A[a * r:(a + 1) * r].t()) @ Bm[off[a]:off[a + 1]].t()
                                   for a in range(n_ad)], 1)
    ref.backward(dkv.float())
    kv = kv0.clone()
    Z = ops.unet_lora_fwd(ehs, A.detach(), Bm.detach(), blk_t, kv, r, 1.5)
    _close(kv, ref.detach(), 2e-3)
    _close(Z, ehs.float() @ A.detach().t(), 1e-5)
    grads = torch.zeros_like(params)
    dA, dB = grads[:A.numel()].view_as(A), grads[A.numel():].view_as(Bm)
    d_ehs = torch.ones(M, ctx)
    ops.unet_lora_bwd(dkv, ehs, A.detach(), Bm.detach(), Z, blk_t, off_t, dA, dB, d_ehs, r, 1.5)
    _close(dA, A.grad, 1e-5)
    _close(dB, Bm.grad, 1e-5)
    _close(d_ehs - 1.0, e32.grad, 1e-5)
    from textboost_b200 import _cabi as C
    with pytest.raises(RuntimeError, match="bad args"):  # rank above the kernels' register budget (16)
        C.call("tb_unet_lora_fwd", C.ptr(ehs), C.ptr(A.detach()), C.ptr(Bm.detach()), C.ptr(blk_t), C.ptr(Z), C.ptr(kv),
               M, ctx, KV, n_ad, 17, 1.0, C.stream_ptr())
    # the optimiser call of the third group: two steps against torch.optim.AdamW on the same gradients
    p_ref = params.clone().requires_grad_(True)
    ref_opt = torch.optim.AdamW([p_ref], lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    ours_p = params.clone()
    opt = FlatAdamW(ours_p, grads, lr=1e-3)
    for it in range(2):
        if it:
            grads.copy_(torch.randn(grads.shape, generator=g))
        p_ref.grad = grads.clone()
        ref_opt.step()
        opt.step()
        assert torch.count_nonzero(grads) == 0
        torch.testing.assert_close(ours_p, p_ref.detach(), rtol=1e-5, atol=1e-7)
    assert opt.state[0].item() == 1.0 and opt.state[4].item() == 2 and opt.state[8].item() == 0


def test_simt_entry_points_under_the_bf16_host_build(monkeypatch):
    """The bf16 precision policy on the CPU: the same SIMT sources built for the host with the 16-bit type set to
    bfloat16 (the host twin of libtextboost_b200_bf16.so, -DTB_BF16) behind the product's ops layer with the policy
    switched to bf16 -- GroupNorm (+SiLU) forward / backward on both paths, LayerNorm, GEGLU, the fp32 -> 16-bit cast,
    noise / MSE, and the UNet-adapter kernels, against fp32 formulas with the bf16 bounds (8x the fp16 ones)."""
    abi = K.install_abi_bf16(monkeypatch)  # noqa: F841
    from textboost_b200 import ops
    from textboost_b200.precision import POLICY
    assert POLICY.act == torch.bfloat16
    BF, TX = torch.bfloat16, 8.0

    def b(*shape, seed=0, scale=1.0, shift=0.0):
        return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale + shift).to(BF)

    # GroupNorm: two-kernel path (cpg 2) and group-owner path (cpg 40)
    for B, HW, Cc, silu in ((2, 40, 64, True), (2, 12, 1280, False)):
        x, gamma, beta = b(B, HW, Cc, seed=Cc, scale=1.5, shift=0.3), b(Cc, seed=1, scale=0.1, shift=1.0), b(Cc, seed=2, scale=0.1)
        y, stats = ops.groupnorm(x, gamma, beta, 32, 1e-5, silu)
        assert y.dtype == BF
        xf = x.float().transpose(1, 2).requires_grad_(True)
        ref = F.group_norm(xf, 32, gamma.float(), beta.float(), 1e-5)
        ref = F.silu(ref) if silu else ref
        _close(y, ref.detach().transpose(1, 2), 2e-3 * TX)
        dy, add = b(B, HW, Cc, seed=3), b(B, HW, Cc, seed=4)
        dx = ops.groupnorm_bwd(dy, x, gamma, beta, stats, 32, 1e-5, silu, add=add)
        ref.backward(dy.float().transpose(1, 2))
        _close(dx, xf.grad.transpose(1, 2) + add.float(), 3e-3 * TX)
    # LayerNorm on 16-bit rows
    x, gamma, beta = b(37, 320, seed=5), b(320, seed=6, scale=0.1, shift=1.0), b(320, seed=7, scale=0.1)
    y, st = ops.layernorm(x, gamma, beta)
    xr = x.float().requires_grad_(True)
    yr = F.layer_norm(xr, (320,), gamma.float(), beta.float(), 1e-5)
    _close(y, yr.detach(), 2e-3 * TX)
    dy = b(37, 320, seed=8)
    yr.backward(dy.float())
    _close(ops.layernorm_bwd(dy, x, gamma, st), xr.grad, 3e-3 * TX)
    # GEGLU and the fp32 -> 16-bit cast
    h = b(9, 64, seed=9)
    hr = h.float().requires_grad_(True)
    gr = hr[:, :32] * F.gelu(hr[:, 32:])
    _close(ops.geglu(h), gr.detach(), 2e-3 * TX)
    dg = b(9, 32, seed=10)
    gr.backward(dg.float())
    _close(ops.geglu_bwd(dg, h), hr.grad, 3e-3 * TX)
    f = torch.randn(5, 16, generator=torch.Generator().manual_seed(11))
    assert torch.equal(ops.cast_f32_f16(f, scale=0.5), (f * 0.5).to(BF))
    # add_noise + MSE (fp32 in, 16-bit noisy sample / prediction)
    acp = torch.linspace(0.999, 0.01, 1000)
    x0, eps = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(12)), torch.randn(2, 4, 8, 8)
    t = torch.tensor([10, 900])
    noisy, target = ops.add_noise(x0, eps, t, acp, False)
    sa, sb = acp[t].sqrt().view(2, 1, 1, 1), (1 - acp[t]).sqrt().view(2, 1, 1, 1)
    assert noisy.dtype == BF and torch.equal(target, eps)
    _close(noisy, sa * x0 + sb * eps, 1e-3 * TX)
    loss = torch.zeros(1)
    dpred = ops.mse_fwd_bwd(noisy, target, loss, 1.0, torch.ones(1))
    _close(loss, F.mse_loss(noisy.float(), target).view(1), 1e-5)
    _close(dpred, 2 * (noisy.float() - target) / noisy.numel(), 1e-3 * TX)
    # the UNet-adapter kernels on bf16 text states / K-V
    r, M, ctx, widths = 4, 6, 40, (8, 16)
    KV, n_ad = 2 * sum(widths), 2 * len(widths)
    blk, off = [], [0]
    for i, w in enumerate(widths):
        blk += [2 * i] * w + [2 * i + 1] * w
        off += [off[-1] + w, off[-1] + 2 * w]
    blk_t, off_t = torch.tensor(blk, dtype=torch.int32), torch.tensor(off, dtype=torch.int32)
    g = torch.Generator().manual_seed(13)
    ehs, kv0, dkv = b(M, ctx, seed=14), b(M, KV, seed=15), b(M, KV, seed=16)
    A = (torch.randn(n_ad * r, ctx, generator=g) / r).requires_grad_(True)
    Bm = (0.05 * torch.randn(KV, r, generator=g)).requires_grad_(True)
    e32 = ehs.float().requires_grad_(True)
    ref = kv0.float() + torch.cat([(e32 @ A[a * r:(a + 1) * r].t()) @ Bm[off[a]:off[a + 1]].t() for a in range(n_ad)], 1)
    ref.backward(dkv.float())
    kv = kv0.clone()
    Z = ops.unet_lora_fwd(ehs, A.detach(), Bm.detach(), blk_t, kv, r, 1.0)
    _close(kv, ref.detach(), 2e-3 * TX)
    dA, dB, d_ehs = torch.zeros_like(A), torch.zeros_like(Bm), torch.zeros(M, ctx)
    ops.unet_lora_bwd(dkv, ehs, A.detach(), Bm.detach(), Z, blk_t, off_t, dA, dB, d_ehs, r, 1.0)
    _close(dA, A.grad, 1e-5)
    _close(dB, Bm.grad, 1e-5)
    _close(d_ehs, e32.grad, 1e-5)
