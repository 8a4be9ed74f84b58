#!/usr/bin/env python3
"""inference.py — the reference's sampling CLI (/root/reference/inference.py) over the B200 sampler.

Same positional argument and flags (``path``, ``--model``, ``--prompt``, ``--outdir``, ``--checkpoint``,
``--skip-gen``, ``--seeds``, ``--output``) and the same steps as the reference's ``load_pipeline`` / ``main``
(:46-113): load the base Stable Diffusion pipeline, ``text_encoder.load_adapter(<path>/text_encoder)``, one
``load_textual_inversion`` per ``*.bin`` except optimizer / scheduler state, swap in
``DPMSolverMultistepScheduler.from_config(pipeline.scheduler.config)``, one generator per seed,
``pipeline(prompt, num_images_per_prompt=len(seeds), generator=[...]).images``, save a grid (``--output``) or one JPEG
per seed.  ``--model`` also accepts a local diffusers-layout directory (there is no hub access here);
``--num_inference_steps`` / ``--guidance_scale`` are additions with the pipeline's defaults (50, 7.5).
"""
import argparse
import os

import torch
from PIL import Image

STABLE_DIFFUSION = {
    "sd14": "CompVis/stable-diffusion-v1-4",
    "sd15": "stable-diffusion-v1-5/stable-diffusion-v1-5",
    "sd21base": "stabilityai/stable-diffusion-2-1-base",
    "sd21": "stabilityai/stable-diffusion-2-1",
}


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("path", type=str, help="path to model")
    parser.add_argument("--model", type=str, default="sd21base")
    parser.add_argument("--prompt", type=str, default="photo of a <dog> dog",
                        help="[sks SUBJECT] for DreamBooth models, [<INSTANCE>] for Textual Inversion models, "
                             "[<INSTANCE> SUBJECT] for CustomDiffusion and TextBoost.")
    parser.add_argument("--outdir", type=str, default="./benchmarks")
    parser.add_argument("--checkpoint", type=int, default=None)
    parser.add_argument("--skip-gen", action="store_true")
    parser.add_argument("--seeds", type=int, nargs="+", default=[0, 1, 2, 3])
    parser.add_argument("--output", type=str, default=None)
    parser.add_argument("--num_inference_steps", type=int, default=50)
    parser.add_argument("--guidance_scale", type=float, default=7.5)
    args = parser.parse_args(argv)
    if args.model in STABLE_DIFFUSION and not os.path.isdir(args.model):
        args.model = STABLE_DIFFUSION[args.model]
    return args


def make_image_grid(images, rows, cols):
    """diffusers.utils.make_image_grid: rows x cols sheet of equally sized images."""
    assert len(images) == rows * cols
    w, h = images[0].size
    grid = Image.new("RGB", size=(cols * w, rows * h))
    for i, img in enumerate(images):
        grid.paste(img, box=(i % cols * w, i // cols * h))
    return grid


def load_pipeline(model_path, pretrained_model, dtype=torch.float16):
    from textboost_b200.pipeline import DiffusionPipeline
    if not os.path.isdir(pretrained_model):
        raise OSError(f"base model {pretrained_model!r} is not a local directory (no hub access): pass --model <dir> "
                      "with the diffusers layout (unet/, vae/, text_encoder/, tokenizer/, scheduler/)")
    pipeline = DiffusionPipeline.from_pretrained(pretrained_model, use_safetensors=True, safety_checker=None)
    text_encoder_path = os.path.join(model_path, "text_encoder")
    pipeline.text_encoder.load_adapter(text_encoder_path, "default")
    print("Loaded text encoder LoRA weights")
    pipeline.text_encoder.set_adapter("default")
    for embedding in sorted(f for f in os.listdir(model_path) if f.endswith(".bin")):
        if os.path.basename(embedding) in ("optimizer.bin", "scheduler.bin"):
            continue
        emb_path = os.path.join(model_path, embedding)
        pipeline.load_textual_inversion(emb_path)
        print(f"Loaded learned embeddings from {emb_path}")
    pipeline.set_progress_bar_config(disable=True)
    pipeline.vae.eval().requires_grad_(False)
    pipeline.unet.eval().requires_grad_(False)
    pipeline.text_encoder.eval().requires_grad_(False)
    return pipeline.to(dtype=dtype)


@torch.inference_mode()
def main(args):
    from textboost_b200.pipeline import DPMSolverMultistepScheduler
    if args.path.endswith("/"):
        args.path = args.path[:-1]
    if not torch.cuda.is_available():
        raise SystemExit("inference.py needs a B200: the CUDA library is the product, there is no CPU path")
    device = torch.device("cuda")
    pipeline = load_pipeline(args.path, args.model)
    pipeline.scheduler = DPMSolverMultistepScheduler.from_config(pipeline.scheduler.config)
    pipeline = pipeline.to(device)
    generator = [torch.Generator(device).manual_seed(seed) for seed in args.seeds]
    images = pipeline(prompt=args.prompt, num_images_per_prompt=len(generator), generator=generator,
                      num_inference_steps=args.num_inference_steps, guidance_scale=args.guidance_scale).images
    if args.output is not None:
        make_image_grid(images, 1, len(args.seeds)).save(args.output)
        return [args.output]
    outputs = []
    for seed, image in zip(args.seeds, images):
        output = args.prompt.replace(" ", "_") + f"_{seed}.jpg"
        image.save(output)
        outputs.append(output)
    return outputs


if __name__ == "__main__":
    main(parse_args())
