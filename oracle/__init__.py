"""oracle/ — CPU restatement of the reference TextBoost training step.  TEST INFRASTRUCTURE ONLY.

Nothing under this package is part of the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it, and only as the checker
(or as the timed CPU baseline), never as something the shipped path routes through.

What it restates (plain PyTorch, fp32/fp64 on CPU, autograd supplies every backward):
  * ddpm_ref.py   — diffusers 0.29.0 DDPMScheduler.add_noise / get_velocity, training_utils.compute_snr,
                    and the timestep sampler of train_textboost.py:991-997.
  * clip_ref.py   — transformers CLIPTextModel (modeling_clip.py), peft 0.13.2 LoRA Linear,
                    textboost/text_encoder.py:17-87 (TextBoostModel null-embedding override).
  * unet_ref.py   — diffusers 0.29.0 UNet2DConditionModel (SD-1.x and SD-2.x switches).
  * step_ref.py   — train_textboost.py:1041-1149, one training step (incl. the --with_image_prior two-part loss).
  * vae_ref.py    — diffusers 0.29.0 AutoencoderKL, encoder + decoder (train_textboost.py:1036-1037; the sampler's
                    decode).
  * sampler_ref.py — diffusers DPMSolverMultistepScheduler at its defaults + the classifier-free-guidance sampling loop
                    of StableDiffusionPipeline (train_textboost.py:453-531, inference.py:84-105).
  * pil_resample_ref.py, pil_affine_ref.py — Pillow's 8-bit antialiased resize, AFFINE transform (bicubic / nearest),
                    grayscale, and torchvision's pad / center_crop index rules: the image arithmetic under
                    textboost/dataset.py:326-351 and textboost/augment/paired_augmentation.py.

Parity pinning: the reference repository ships NO tests, golden vectors or fixtures for this path
(SURVEY.md §4, §8c), and diffusers / peft / accelerate are not installable here.  What *is* pinned:
the CLIP + TextBoostModel part is checked bit-for-bit (fp32) against the reference's own
``textboost.text_encoder.TextBoostModel`` running on the installed transformers 5.5.0
(tests/golden/ fixtures made by tests/golden/make_golden.py), AdamW / grad clipping against torch's own
implementations, and DDPM constants against their closed forms.  The UNet restatement follows the
published diffusers 0.29.0 architecture and has no executable reference here:
**UNet parity is unpinned** (stated again in DESIGN.md); likewise vae_ref and sampler_ref (checked by published
parameter counts, the known sigma schedule and the solver's fixed-point property only).  The two pil_* modules ARE
pinned: byte for byte against the installed Pillow / torchvision (tests/test_resample_cpu.py).
tests/test_oracle_vs_libraries_cpu.py cross-checks every unpinned module against diffusers / peft wherever they import.
"""
