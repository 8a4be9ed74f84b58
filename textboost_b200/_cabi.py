"""ctypes binding of libtextboost_b200.so (declared in include/textboost_b200.h).

The library is the product: there is NO fallback.  Loading fails loudly when the shared object is
missing (run ``python -c 'import __graft_entry__ as g; g.build()'``), and every compute entry point
returns TB_E_ARCH on a device that is not sm_100, which `call()` turns into a RuntimeError.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_void_p)

from .precision import POLICY

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path() -> str:
    """The build of the library that matches the process's precision policy (precision.py).  TB_LIB points at an
    alternative build of the same library (A/B timing of kernel variants); never a fallback."""
    return os.environ.get("TB_LIB") or os.path.join(_HERE, "lib", POLICY.lib_name)

TB_ACT_NONE, TB_ACT_SILU, TB_ACT_QUICK_GELU, TB_ACT_GELU = 0, 1, 2, 3
TB_OUT_F16, TB_OUT_F32, TB_OUT_F32_ACC = 0, 1, 2


class Epilogue(Structure):
    _fields_ = [
        ("bias", c_void_p),
        ("rowvec", c_void_p),
        ("rows_per_group", c_int32),
        ("residual", c_void_p),
        ("ldr", c_int64),
        ("residual_f32", c_int32),
        ("alpha", c_float),
        ("act", c_int32),
        ("out_kind", c_int32),
        ("ld_rowvec", c_int64),
    ]


_lib = None

# name -> argtypes (restype is always int unless listed in _RESTYPES)
_SIGNATURES = {
    "tb_version": [],
    "tb_last_error": [],
    "tb_check_device": [],
    "tb_storage_dtype": [],
    "tb_set_workspace": [c_void_p, c_void_p, c_size_t],
    "tb_gemm_f16": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                    POINTER(Epilogue), c_void_p],
    "tb_conv3x3_f16": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                       POINTER(Epilogue), c_void_p],
    "tb_attn_fwd_f16": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                        c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p],
    "tb_attn_bwd_f16": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                        c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                        c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p],
    "tb_attn_debug_trace": [c_void_p],
    "tb_groupnorm_fwd_f16": [c_void_p] * 5 + [c_int] * 4 + [c_float, c_int, c_void_p],
    "tb_groupnorm_bwd_f16": [c_void_p] * 8 + [c_int] * 4 + [c_float, c_int, c_void_p],
    "tb_layernorm_fwd": [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int64,
                         c_void_p, c_int, c_int, c_float, c_void_p],
    "tb_layernorm_lora_fwd": [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int,
                              c_int, c_int, c_float, c_void_p],
    "tb_layernorm_bwd_clip": [c_void_p, c_int, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "tb_layernorm_bwd": [c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p,
                         c_void_p, c_int, c_int, c_void_p],
    "tb_geglu_fwd_f16": [c_void_p, c_void_p, c_int64, c_int, c_void_p],
    "tb_geglu_bwd_f16": [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p],
    "tb_upsample2x_fwd_f16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "tb_upsample2x_bwd_f16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "tb_concat2_f16": [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int64, c_void_p],
    "tb_copy2d_f16": [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p],
    "tb_cast_f32_f16": [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_float, c_void_p],
    "tb_im2col3x3s2_f16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "tb_zero_stuff2x_f16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "tb_im2col3x3s2_pad_f16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "tb_softmax_rows_f16": [c_void_p, c_int64, c_int64, c_int, c_void_p],
    "tb_vae_sample": [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                      c_void_p],
    "tb_dpm_cfg_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_int,
                        c_float, c_float, c_float, c_void_p],
    "tb_vae_decode_in": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p],
    "tb_image_u8": [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p],
    "tb_img_gather_u8": [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                         c_int, c_void_p],
    "tb_img_affine_u8": [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p],
    "tb_img_grayscale_u8": [c_void_p, c_void_p, c_int64, c_void_p],
    "tb_resize_crop_normalize_u8": [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                    c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                    c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p],
    "tb_timestep_embedding_f16": [c_void_p, c_void_p, c_int, c_int, c_void_p],
    "tb_silu_f16": [c_void_p, c_void_p, c_int64, c_void_p],
    "tb_axpy_f32": [c_void_p, c_void_p, c_int64, c_float, c_void_p],
    "tb_fill_zero": [c_void_p, c_size_t, c_void_p],
    "tb_add_noise": [c_void_p] * 6 + [c_int, c_int, c_int, c_void_p],
    "tb_mse_fwd_bwd": [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p],
    "tb_conv_in_f16": [c_void_p] * 4 + [c_int] * 5 + [c_void_p],
    "tb_conv_out_f16": [c_void_p] * 4 + [c_int] * 5 + [c_void_p],
    "tb_conv_out_bwd_f16": [c_void_p] * 3 + [c_int] * 5 + [c_void_p],
    "tb_clip_embed": [c_void_p] * 6 + [c_int] * 5 + [c_void_p],
    "tb_clip_embed_grad": [c_void_p] * 3 + [c_int] * 3 + [c_void_p],
    "tb_lora_down": [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "tb_lora_pack": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p],
    "tb_lora_grad": [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                     c_float, c_void_p],
    "tb_lora_dx": [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p],
    "tb_unet_lora_fwd": [c_void_p] * 6 + [c_int] * 5 + [c_float, c_void_p],
    "tb_unet_lora_bwd": [c_void_p] * 11 + [c_int] * 5 + [c_float, c_void_p],
    "tb_clip_attn_fwd": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "tb_clip_attn_bwd": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "tb_act_fwd_f16": [c_void_p, c_void_p, c_int64, c_int, c_void_p],
    "tb_act_bwd_f16": [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p],
    "tb_null_override": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "tb_kpl_fwd_bwd": [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                       c_void_p],
    "tb_optim_mix_mask": [c_void_p, c_int64, c_int, c_int, c_int, c_void_p],
    "tb_adamw_fused_step": [c_void_p] * 4 + [c_int64, c_int, c_int] + [c_float] * 9 + [c_int, c_float, c_float] +
                           [c_void_p, c_void_p, c_void_p],
}
_RESTYPES = {"tb_last_error": c_char_p}


def exported_symbols():
    """Names the header declares (used by the CPU-side ABI test)."""
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: the CUDA extension is the product and has no fallback. "
                "Build it with `python -m textboost_b200.build`.")
        l = ctypes.CDLL(path)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        if l.tb_storage_dtype() != POLICY.storage_code:
            raise RuntimeError(f"{path} was built for storage type {l.tb_storage_dtype()} but the process's precision "
                               f"policy is {POLICY.name} (expects {POLICY.storage_code})")
        POLICY.locked = True
        _lib = l
    return _lib


def last_error() -> str:
    return lib().tb_last_error().decode(errors="replace")


# KERNELS one call of each entry point enqueues (memset nodes are not counted); entry points not listed launch one
_KERNELS_PER_CALL = {"tb_groupnorm_fwd_f16": 2, "tb_groupnorm_bwd_f16": 2, "tb_attn_bwd_f16": 2, "tb_unet_lora_fwd": 2, "tb_unet_lora_bwd": 4,
                     "tb_resize_crop_normalize_u8": 2,
                     "tb_adamw_fused_step": 3, "tb_fill_zero": 0, "tb_version": 0, "tb_check_device": 0, "tb_storage_dtype": 0, "tb_set_workspace": 0,
                     "tb_attn_debug_trace": 0}
TB_GN_STATS_ZEROED = 2
launch_count = 0  # GPU launches enqueued through this binding since import (bench.py reports the delta)


def call(name: str, *args):
    global launch_count
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {last_error()}")
    launch_count += _KERNELS_PER_CALL.get(name, 1)


WORKSPACE_BYTES = 32 << 20
_workspaces = {}  # (device index, stream handle) -> tensor kept alive for the library


def stream_ptr():
    """Handle of torch's current stream.  The first use of a stream registers a split-K workspace for it
    (tb_set_workspace): the library never allocates, PyTorch owns the memory."""
    import torch
    s = torch.cuda.current_stream()
    key = (s.device.index, s.cuda_stream)
    if key not in _workspaces:
        buf = torch.empty(WORKSPACE_BYTES, device=s.device, dtype=torch.uint8)
        _workspaces[key] = buf
        call("tb_set_workspace", c_void_p(s.cuda_stream), c_void_p(buf.data_ptr()), WORKSPACE_BYTES)
    return c_void_p(s.cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())
