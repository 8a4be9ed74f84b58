"""The bf16 precision policy (--mixed_precision bf16, train_textboost.py:932-933): libtextboost_b200_bf16.so, the second
build of the same kernel sources, goes through the SAME kernel-level and path-level parity tests as the fp16 build.

A process holds one policy (textboost_b200/precision.py), so the two files are re-run in a child interpreter with
TEXTBOOST_B200_PRECISION=bf16: there every 16-bit tensor is torch.bfloat16, the library loaded is the bf16 one (the
loader checks tb_storage_dtype()), and every fp16 bound is allowed the 8x of bf16's unit roundoff
(conftest.tol_scale).  The CLI is driven with --mixed_precision bf16 in a third child.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(built_lib):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    if os.environ.get("TEXTBOOST_B200_PRECISION", "fp16") != "fp16":
        pytest.skip("already inside the bf16 child run")


def _child(args, log_name, timeout=900):
    env = dict(os.environ, TEXTBOOST_B200_PRECISION="bf16")
    r = subprocess.run([sys.executable] + args, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", log_name), "w") as f:
        f.write(r.stdout[-20000:] + "\n--- stderr ---\n" + r.stderr[-5000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    return r.stdout


def test_kernel_suite_under_the_bf16_build():
    out = _child(["-m", "pytest", "tests/test_gpu_kernels.py", "-m", "gpu", "-q", "-p", "no:cacheprovider"],
                 "pytest_bf16_kernels.log")
    assert " passed" in out and " failed" not in out


def test_step_suite_under_the_bf16_build():
    out = _child(["-m", "pytest", "tests/test_gpu_step.py", "-m", "gpu", "-q", "-p", "no:cacheprovider"],
                 "pytest_bf16_step.log")
    assert " passed" in out and " failed" not in out


def test_vae_and_sampler_suites_under_the_bf16_build():
    """The front end (AutoencoderKL encoder) and the sampler (DPM-Solver++ / DDPM loops, VAE decoder) on the bf16 build;
    the CLI cases of those files pass --mixed_precision fp16 and stay with the fp16 run."""
    out = _child(["-m", "pytest", "tests/test_gpu_vae.py", "tests/test_gpu_sampler.py", "-m", "gpu", "-q", "-k",
                  "not cli", "-p", "no:cacheprovider"], "pytest_bf16_vae_sampler.log")
    assert " passed" in out and " failed" not in out


def test_cli_mixed_precision_bf16(tmp_path):
    """train_textboost.py --mixed_precision bf16 in a fresh interpreter: the CLI itself selects the bf16 library (no
    environment variable), trains without a GradScaler (loss scale 1, nothing skipped) and writes the output files."""
    code = f"""
import os, sys, torch
sys.path.insert(0, {ROOT!r})
os.environ.pop("TEXTBOOST_B200_PRECISION", None)
import train_textboost as T
from textboost_b200 import synthetic, _cabi, precision
ck, out = {str(tmp_path / 'model')!r}, {str(tmp_path / 'out')!r}
synthetic.write_pretrained(ck, "tiny", seed=2)
loss = T.main(T.parse_args(["--pretrained_model_name_or_path", ck, "--output_dir", out, "--synthetic_data",
    "--resolution", "128", "--train_batch_size", "2", "--learning_rate", "1e-3", "--mixed_precision", "bf16",
    "--max_train_steps", "8", "--log_every", "4"]))
assert loss == loss and loss < 10, loss
assert precision.POLICY.name == "bf16" and _cabi.lib().tb_storage_dtype() == 1
assert _cabi.lib_path().endswith("libtextboost_b200_bf16.so")
assert {{"text_encoder", "dog.bin"}} <= set(os.listdir(out))
assert T.RUN_INFO["loss_scale"] == 1.0 and T.RUN_INFO["skipped_steps"] == 0, T.RUN_INFO  # no GradScaler under bf16
# --unet_params_to_train crossattn_kv: runs under bf16 (the reference's fp16 path cannot), writes the UNet adapter
out2 = out + "_kv"
loss2 = T.main(T.parse_args(["--pretrained_model_name_or_path", ck, "--output_dir", out2, "--synthetic_data",
    "--resolution", "128", "--train_batch_size", "2", "--learning_rate", "1e-3", "--mixed_precision", "bf16",
    "--max_train_steps", "8", "--checkpointing_steps", "4", "--unet_params_to_train", "crossattn_kv"]))
assert loss2 == loss2 and loss2 < 10, loss2
from safetensors.torch import load_file
ad = load_file(os.path.join(out2, "unet", "pytorch_lora_weights.safetensors"))
kb = [k for k in ad if k.endswith("attn2.to_v.lora_B.weight")]
assert len(kb) == len(ad) // 4 and all(k.startswith("unet.") for k in ad)
assert sum(float(ad[k].abs().sum()) for k in kb) > 0  # lora_B left its zero init: the UNet adapter trained
assert os.path.exists(os.path.join(out2, "checkpoint-4", "pytorch_lora_weights.safetensors"))
loss3 = T.main(T.parse_args(["--pretrained_model_name_or_path", ck, "--output_dir", out2, "--synthetic_data",
    "--resolution", "128", "--train_batch_size", "2", "--learning_rate", "1e-3", "--mixed_precision", "bf16",
    "--max_train_steps", "10", "--checkpointing_steps", "4", "--unet_params_to_train", "crossattn_kv",
    "--resume_from_checkpoint", "latest"]))
assert loss3 == loss3
print("BF16_CLI_OK", loss, loss2, loss3)
"""
    env = {k: v for k, v in os.environ.items() if k != "TEXTBOOST_B200_PRECISION"}
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "BF16_CLI_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
