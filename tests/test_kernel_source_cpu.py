"""The source of the fp16 elementwise kernels of csrc/vae.cu and csrc/sampler.cu, compiled for the host
(tests/kernel_host_emulation.py: g++, __half = _Float16, one emulated thread of the grid-stride loops) and checked on
the CPU against the formulas the GPU tests use — regression cover, in the CPU suite, for kernels whose hardware parity
was measured once (tests/test_gpu_vae.py, tests/test_gpu_sampler.py)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import kernel_host_emulation as K
import ops_standin


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize("pad_lo", [0, 1])
def test_im2col_stride2_kernel_source(pad_lo):
    x = torch.randn(2, 12, 16, 16).half()
    col = torch.empty(2 * 6 * 8, 9 * 16, dtype=torch.float16)
    K.lib_f16().emu_im2col_pad(_p(x), _p(col), 2, 12, 16, 16, pad_lo)
    assert torch.equal(col, ops_standin.im2col3x3s2_pad(x, pad_lo))


def test_vae_sample_kernel_source():
    g = torch.Generator().manual_seed(1)
    B, HW, L = 3, 40, 4
    rows = torch.randn(B * HW, 64, generator=g)
    rows[:, L:2 * L] *= 20
    rows = rows.half()
    eps = torch.randn(B, L, HW, generator=g)
    lat, mean, std = (torch.empty(B, L, HW) for _ in range(3))
    K.lib_f16().emu_vae_sample(_p(rows), ctypes.c_longlong(64), _p(eps), _p(lat), _p(mean), _p(std), HW, L,
                               ctypes.c_longlong(B * L * HW), ctypes.c_float(0.18215))
    lat_r, mean_r, std_r = ops_standin.vae_sample(rows, B, HW, L, eps=eps, scaling_factor=0.18215, want_moments=True)
    assert torch.equal(mean, mean_r)
    torch.testing.assert_close(std, std_r, rtol=2e-7, atol=0)
    torch.testing.assert_close(lat, lat_r, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("v_pred", [0, 1])
def test_dpm_cfg_step_kernel_source_over_a_whole_schedule(v_pred):
    from textboost_b200.pipeline import DPMSolverMultistepScheduler
    s = DPMSolverMultistepScheduler(prediction_type="v_prediction" if v_pred else "epsilon")
    s.set_timesteps(7)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 4, 8, 8, generator=g) * 10
    xr = x.clone()
    n = x.numel()
    m, mr = [torch.zeros_like(x) for _ in range(2)], [torch.zeros_like(x) for _ in range(2)]
    for i in range(7):
        eps = torch.randn(4, 4, 8, 8, generator=g).half()
        k = s.step_coefficients(i)
        uin, uin_r = torch.zeros(4, 4, 8, 8, dtype=torch.float16), torch.zeros(4, 4, 8, 8, dtype=torch.float16)
        prev, prev_r = (m[(i + 1) % 2], mr[(i + 1) % 2]) if k["c_d1"] else (None, None)
        K.lib_f16().emu_dpm_cfg_step(_p(x), _p(eps), _p(prev) if prev is not None else None, _p(m[i % 2]), _p(uin),
                                     ctypes.c_longlong(n), ctypes.c_float(7.5), ctypes.c_float(k["alpha_i"]),
                                     ctypes.c_float(k["sigma_i"]), v_pred, ctypes.c_float(k["c_x"]),
                                     ctypes.c_float(k["c_d0"]), ctypes.c_float(k["c_d1"]))
        ops_standin.dpm_cfg_step(xr, eps, prev_r, mr[i % 2], uin_r, 7.5, k["alpha_i"], k["sigma_i"], bool(v_pred),
                                 k["c_x"], k["c_d0"], k["c_d1"])
        torch.testing.assert_close(x, xr, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(m[i % 2], mr[i % 2], rtol=1e-5, atol=1e-5)
        assert torch.equal(uin[:2], uin[2:]) and (uin.float() - uin_r.float()).abs().max() <= 0.02
    torch.testing.assert_close(x, mr[6 % 2], rtol=1e-5, atol=1e-5)  # the last step returns the data prediction


def test_vae_decode_in_and_image_u8_kernel_source():
    g = torch.Generator().manual_seed(3)
    lat = torch.randn(2, 4, 5, 7, generator=g)
    w, b = torch.randn(4, 4, generator=g) * 0.5, torch.randn(4, generator=g) * 0.1
    z = torch.empty(2, 4, 5, 7, dtype=torch.float16)
    K.lib_f16().emu_vae_decode_in(_p(lat), _p(w), _p(b), _p(z), 4, 35, ctypes.c_longlong(2 * 4 * 35),
                                  ctypes.c_float(1.0 / 0.18215))
    zr = ops_standin.vae_decode_in(lat, w, b, 0.18215)
    assert (z.float() - zr.float()).abs().max() <= 2e-3 * zr.float().abs().max()
    rows = (torch.randn(300, 64, generator=g) * 0.8).half()
    rows[:3, :3] = torch.tensor([[-1.0, 1.0, 0.0], [-3.0, 3.0, 0.5], [1 / 255 - 1, 3 / 255 - 1, 0.25]])
    u8 = torch.empty(300, 3, dtype=torch.uint8)
    K.lib_f16().emu_image_u8(_p(rows), ctypes.c_longlong(64), _p(u8), ctypes.c_longlong(300), 3)
    assert torch.equal(u8, ops_standin.image_u8(rows, 300, 3))
    assert u8[0].tolist() == [0, 255, 128] and u8[1].tolist()[:2] == [0, 255]
