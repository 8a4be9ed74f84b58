#!/usr/bin/env python3
"""inference.py — the reference's sampling CLI (/root/reference/inference.py) over the B200 sampler.

Same command line: positional ``path`` (a TextBoost output directory) and ``--model``, ``--prompt``, ``--outdir``,
``--checkpoint``, ``--skip-gen``, ``--seeds``, ``--output`` with the reference's types and defaults (pinned by
tests/golden/inference_flags.json); ``--num_inference_steps`` / ``--guidance_scale`` are additions carrying the
pipeline defaults (50, 7.5).  Same steps as the reference's ``load_pipeline`` / ``main`` (:46-113): base Stable
Diffusion pipeline -> ``text_encoder.load_adapter(<path>/text_encoder)`` -> one ``load_textual_inversion`` per learned
``*.bin`` (optimizer / scheduler state files skipped) -> ``DPMSolverMultistepScheduler.from_config`` -> one generator
per seed -> ``pipeline(prompt, num_images_per_prompt=len(seeds), generator=[...]).images`` -> a one-row grid
(``--output``) or ``<prompt with underscores>_<seed>.jpg`` per seed.  There is no hub access here, so ``--model`` must
resolve to a local directory in the diffusers layout; the reference's short names are still mapped to their hub ids.
"""
import argparse
import os

import torch
from PIL import Image

STABLE_DIFFUSION = dict(sd14="CompVis/stable-diffusion-v1-4", sd15="stable-diffusion-v1-5/stable-diffusion-v1-5",
                        sd21base="stabilityai/stable-diffusion-2-1-base", sd21="stabilityai/stable-diffusion-2-1")
_NOT_EMBEDDINGS = {"optimizer.bin", "scheduler.bin"}  # accelerate state that may sit next to the learned rows

_FLAGS = [
    ("path", dict(type=str, help="path to model")),
    ("--model", dict(type=str, default="sd21base")),
    ("--prompt", dict(type=str, default="photo of a <dog> dog",
                      help="[sks SUBJECT] for DreamBooth models, [<INSTANCE>] for Textual Inversion models, "
                           "[<INSTANCE> SUBJECT] for CustomDiffusion and TextBoost.")),
    ("--outdir", dict(type=str, default="./benchmarks")),
    ("--checkpoint", dict(type=int, default=None)),
    ("--skip-gen", dict(action="store_true")),
    ("--seeds", dict(type=int, nargs="+", default=[0, 1, 2, 3])),
    ("--output", dict(type=str, default=None)),
    ("--num_inference_steps", dict(type=int, default=50)),   # addition
    ("--guidance_scale", dict(type=float, default=7.5)),      # addition
]


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Sample images from a TextBoost run on a B200.")
    for name, spec in _FLAGS:
        parser.add_argument(name, **spec)
    args = parser.parse_args(argv)
    if not os.path.isdir(args.model):
        args.model = STABLE_DIFFUSION.get(args.model, args.model)
    return args


def make_image_grid(images, rows, cols):
    """rows x cols sheet of equally sized images, filled row by row (diffusers.utils.make_image_grid)."""
    if len(images) != rows * cols:
        raise ValueError(f"{len(images)} images do not fill a {rows} x {cols} grid")
    w, h = images[0].size
    sheet = Image.new("RGB", size=(cols * w, rows * h))
    for k, img in enumerate(images):
        sheet.paste(img, box=((k % cols) * w, (k // cols) * h))
    return sheet


def learned_embedding_files(model_path):
    """The ``{token}.bin`` files of a run directory, in name order."""
    return [os.path.join(model_path, f) for f in sorted(os.listdir(model_path))
            if f.endswith(".bin") and f not in _NOT_EMBEDDINGS]


def load_pipeline(model_path, pretrained_model, dtype=None):
    from textboost_b200.pipeline import DiffusionPipeline
    from textboost_b200.precision import POLICY
    dtype = dtype or POLICY.act  # the reference passes torch.float16 (inference.py:41); bf16 under TEXTBOOST_B200_PRECISION=bf16
    if not os.path.isdir(pretrained_model):
        raise OSError(f"base model {pretrained_model!r} is not a local directory (no hub access): pass --model <dir> "
                      "with the diffusers layout (unet/, vae/, text_encoder/, tokenizer/, scheduler/)")
    pipe = DiffusionPipeline.from_pretrained(pretrained_model, use_safetensors=True, safety_checker=None)
    pipe.text_encoder.load_adapter(os.path.join(model_path, "text_encoder"), "default")
    pipe.text_encoder.set_adapter("default")
    print("Loaded text encoder LoRA weights")
    for emb_path in learned_embedding_files(model_path):
        pipe.load_textual_inversion(emb_path)
        print(f"Loaded learned embeddings from {emb_path}")
    pipe.set_progress_bar_config(disable=True)
    for module in (pipe.vae, pipe.unet, pipe.text_encoder):
        module.eval().requires_grad_(False)
    return pipe.to(dtype=dtype)


@torch.inference_mode()
def main(args):
    from textboost_b200.pipeline import DPMSolverMultistepScheduler
    if not torch.cuda.is_available():
        raise SystemExit("inference.py needs a B200: the CUDA library is the product, there is no CPU path")
    device = torch.device("cuda")
    pipe = load_pipeline(args.path.rstrip("/"), args.model)
    pipe.scheduler = DPMSolverMultistepScheduler.from_config(pipe.scheduler.config)
    pipe = pipe.to(device)
    generators = [torch.Generator(device).manual_seed(seed) for seed in args.seeds]
    images = pipe(prompt=args.prompt, num_images_per_prompt=len(generators), generator=generators,
                  num_inference_steps=args.num_inference_steps, guidance_scale=args.guidance_scale).images
    if args.output is not None:
        make_image_grid(images, 1, len(images)).save(args.output)
        return [args.output]
    stem = args.prompt.replace(" ", "_")
    written = []
    for seed, image in zip(args.seeds, images):
        written.append(f"{stem}_{seed}.jpg")
        image.save(written[-1])
    return written


if __name__ == "__main__":
    main(parse_args())
