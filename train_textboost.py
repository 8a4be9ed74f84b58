#!/usr/bin/env python
"""train_textboost.py — the reference CLI (/root/reference/train_textboost.py) over the B200 TextBoost step.

Every flag of the reference's ``parse_args`` (:49-450) is kept with its name, type and default, including the
quirks (``--disable_weighted_sample`` is store_true with default True, so the SNR-weighted timestep sampler
of :991-997 is unreachable there; ``--no-disable_weighted_sample`` is added here to reach it).  ``main``
follows the reference's set-up order (:598-939) through this package's mirrors of its classes —
``TextBoostModel`` / ``UNet2DConditionModel`` ``.from_pretrained``, ``add_token``,
``add_augmentation_tokens``, ``LoraConfig`` + ``add_adapter`` — and replaces the loop body (:1041-1149) with
``TextBoostTrainer.step`` (one CUDA graph; one all-reduce of the flat LoRA + added-row gradient buffer).
Launch one process per GPU with torchrun, as the reference does (README.md:82).

Outputs are the reference's: ``<output_dir>/text_encoder/adapter_{config.json,model.safetensors}``,
``<output_dir>/<token>.bin`` per added token, ``checkpoint-N/`` directories (here with a working resume).

Data: with ``--instance_data_dir`` / ``--concepts_list`` the reference's front end runs —
``TextBoostDataset`` + ``PairedAugmentation`` (textboost_b200.dataset / .augment, same draws as the reference) ->
``pixel_values`` -> the B200 AutoencoderKL encoder (textboost_b200.vae) -> latents, every step.  ``--latents_file``
(a ``torch.save``d dict with ``latents`` [N,4,h,w] fp32 — already scaled by the VAE factor — ``input_ids`` [N,77]
and optional ``prior_ids`` [P,77]) and ``--synthetic_data`` bypass it; both flags are additions, everything else is
the reference's.  ``--validation_prompts`` samples images every ``--validation_steps`` steps with the B200 sampler
(textboost_b200.pipeline, §8 f3) and writes ``validation_<step>.jpg``.  ``--report_to tensorboard`` (the default) writes
loss / lr / added_embedding_norm scalars and the validation images as event files under ``<output_dir>/logs/textboost``.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import shutil
import time
import warnings

import torch

logger = logging.getLogger("textboost")
RUN_INFO = {}  # facts about the last main() call (tests and notebooks read it; the log file has the same lines)

# (name, kwargs) in the reference's order; help strings omitted on purpose (see the reference for prose)
_FLAGS = [
    ("pretrained_model_name_or_path", dict(type=str, default=None, required=True)),
    ("revision", dict(type=str, default=None)), ("variant", dict(type=str, default=None)),
    ("tokenizer_name", dict(type=str, default=None)), ("instance_data_dir", dict(type=str, default=None)),
    ("instance", dict(type=str)), ("class_data_dir", dict(type=str, default=None)),
    ("instance_token", dict(type=str, default=None)), ("class_token", dict(type=str, nargs="+", default=None)),
    ("with_image_prior", dict(default=False, action="store_true")),
    ("image_ppl_weight", dict(type=float, default=1.0)), ("kpl_weight", dict(type=float, default=0.1)),
    ("kpl_type", dict(type=str, default="cos")), ("num_prior_images", dict(type=int, default=200)),
    ("output_dir", dict(type=str, default="dreambooth-model")), ("seed", dict(type=int, default=42)),
    ("resolution", dict(type=int, default=512)), ("center_crop", dict(default=False, action="store_true")),
    ("train_batch_size", dict(type=int, default=1)), ("sample_batch_size", dict(type=int, default=4)),
    ("max_train_steps", dict(type=int, default=500)), ("checkpointing_steps", dict(type=int, default=100)),
    ("checkpoints_total_limit", dict(type=int, default=None)),
    ("resume_from_checkpoint", dict(type=str, default=None)),
    ("gradient_accumulation_steps", dict(type=int, default=1)),
    ("gradient_checkpointing", dict(action="store_true")),
    ("learning_rate", dict(type=float, default=5e-5)), ("emb_learning_rate", dict(type=float, default=1e-3)),
    ("scale_lr", dict(action="store_true", default=False)), ("lr_scheduler", dict(type=str, default="constant")),
    ("lr_warmup_steps", dict(type=int, default=500)), ("dataloader_num_workers", dict(type=int, default=2)),
    ("adam_beta1", dict(type=float, default=0.9)), ("adam_beta2", dict(type=float, default=0.999)),
    ("adam_weight_decay", dict(type=float, default=1e-2)), ("adam_epsilon", dict(type=float, default=1e-08)),
    ("max_grad_norm", dict(default=1.0, type=float)), ("hub_token", dict(type=str, default=None)),
    ("logging_dir", dict(type=str, default="logs")), ("allow_tf32", dict(action="store_true")),
    ("report_to", dict(type=str, default="tensorboard")),
    ("validation_prompts", dict(type=str, nargs="+", default=None)),
    ("num_validation_images", dict(type=int, default=4)), ("validation_steps", dict(type=int, default=100)),
    ("mixed_precision", dict(type=str, default=None, choices=["no", "fp16", "bf16"])),
    ("prior_generation_precision", dict(type=str, default=None, choices=["no", "fp32", "fp16", "bf16"])),
    ("concepts_list", dict(type=str, default=None)),
    ("text_encoder_use_attention_mask", dict(action="store_true")),
    ("skip_save_text_encoder", dict(action="store_true")),
    ("class_labels_conditioning", dict(default=None)),
    ("validation_scheduler", dict(type=str, default="DPMSolverMultistepScheduler",
                                  choices=["DPMSolverMultistepScheduler", "DDPMScheduler"])),
    ("no_safe_serialization", dict(action="store_true")),
    ("placeholder_token", dict(type=str, default="<dog>")), ("initializer_token", dict(type=str, default="dog")),
    ("unet_params_to_train", dict(type=str, default="none",
                                  choices=["none", "crossattn_kv", "crossattn", "attn", "all"])),
    ("augment", dict(default="none")), ("augment_ops", dict(type=str, default="object")),
    ("augment_p", dict(type=float, default=0.8)), ("augment_prompt", dict(type=int, default=1)),
    ("augment_inversion", dict(action="store_true", default=False)), ("num_samples", dict(type=int, default=None)),
    ("lora_rank", dict(type=int, default=4)),
    ("disable_weighted_sample", dict(action="store_true", default=True)),
    ("null_prob", dict(type=float, default=0.1)), ("template", dict(type=str, default="textboost")),
    ("mixing", dict(action="store_true", default=False)),
]


def parse_args(input_args=None):
    parser = argparse.ArgumentParser(description="TextBoost training on B200 (reference CLI surface).")
    for name, kw in _FLAGS:
        parser.add_argument("--" + name, **kw)
    # additions (not in the reference)
    parser.add_argument("--no-disable_weighted_sample", dest="disable_weighted_sample", action="store_false",
                        help="reach the SNR-weighted timestep sampler of train_textboost.py:991-997")
    parser.add_argument("--latents_file", type=str, default=None,
                        help="torch.save'd {'latents','input_ids'[, 'prior_ids']} replacing the image/VAE front end")
    parser.add_argument("--synthetic_data", action="store_true",
                        help="synthetic latents + literal prompt ids (BASELINE.json workload)")
    parser.add_argument("--null_embedding", type=str, default=None,
                        help="[77, hidden] tensor file; default: assets/null_emb_sd21base.pt when the width "
                             "matches, else the frozen encoder's output for the empty prompt (SURVEY.md D6)")
    parser.add_argument("--prior_prompts_file", type=str, default="data/human-written-prompts.jsonl",
                        help="JSONL of human-written prompts for the knowledge-preservation loss (the path the "
                             "reference hard-codes, train_textboost.py:893); used when the file exists")
    parser.add_argument("--log_every", type=int, default=10, help="host read-back period of the loss scalar")
    parser.add_argument("--gpu_image_transforms", action="store_true",
                        help="finish each image on the GPU: the dataset stops after the PIL augmentation and the "
                             "Lanczos resize / crop / normalise run as byte-exact CUDA kernels (same pixel_values); "
                             "decoded source images are cached")
    parser.add_argument("--gpu_augment", action="store_true",
                        help="also run the augmentation's image operations on the GPU: PairedAugmentation draws its "
                             "random numbers and edits the caption as always but records its image operations, which "
                             "run as exact CUDA kernels on the decoded image resident in HBM (implies "
                             "--gpu_image_transforms)")
    args = parser.parse_args(input_args)
    args.gpu_image_transforms = args.gpu_image_transforms or args.gpu_augment

    # post-parse validation: train_textboost.py:435-448
    if args.with_image_prior:
        if args.class_data_dir is None:
            raise ValueError("You must specify a data directory for class images.")
        if args.class_token is None:
            raise ValueError("You must specify prompt for class images.")
    else:
        if args.class_data_dir is not None:
            warnings.warn("You need not use --class_data_dir without --with_image_prior.")
        if args.class_token is not None:
            warnings.warn("You need not use --class_token without --with_image_prior.")
    if args.augment_inversion and not bool(args.augment_prompt):
        raise ValueError("You need to use --augment_prompt=1 with --augment_prompt.")
    return args


# ------------------------------------------------------------------------------------ helpers
def _unsupported(args):
    """Modes of the reference CLI outside the hot path built here: fail loudly."""
    if args.with_image_prior and (args.latents_file or args.synthetic_data):
        raise ValueError("--with_image_prior takes its class images from --class_data_dir through the image front "
                         "end: it cannot be combined with --latents_file / --synthetic_data")
    if args.unet_params_to_train == "crossattn_kv" and args.lora_rank > 0 and args.mixed_precision != "bf16":
        # train_textboost.py:711-720 adds LoRA to attn2.to_k / to_v, then :937 casts the WHOLE UNet -- adapter included --
        # to weight_dtype.  Under --mixed_precision fp16 the trainable tensors are fp16 and accelerate's
        # clip_grad_norm_ -> GradScaler.unscale_ raises "Attempting to unscale FP16 gradients": the mode only runs in
        # the reference's bf16 policy (no GradScaler: built here) and its fp32 policy (not built).
        raise NotImplementedError("--unet_params_to_train crossattn_kv needs --mixed_precision bf16: the reference's "
                                  "fp16 path fails in GradScaler.unscale_ on the fp16 adapter tensors, and fp32 "
                                  "(--mixed_precision no) is not built")
    if args.lora_rank < 0:
        raise ValueError("--lora_rank must be >= 0")
    if args.gradient_accumulation_steps < 1:
        raise ValueError("--gradient_accumulation_steps must be >= 1")
    if args.mixed_precision == "no":
        # weight_dtype = torch.float32 (train_textboost.py:928): every tensor-core operand would be fp32 / tf32
        raise NotImplementedError("--mixed_precision no (fp32 weights and activations) is not built: the B200 path "
                                  "keeps 16-bit tensor-core operands with fp32 accumulation and fp32 master weights "
                                  "(--mixed_precision fp16 or bf16)")
    if args.text_encoder_use_attention_mask:
        # train_textboost.py:1058 hands encode_prompt the collated attention_mask, which is a Python LIST
        # (dataset.py:428, 455): the reference itself fails on `.to(device)` there
        raise NotImplementedError("--text_encoder_use_attention_mask: the reference's own path fails with this flag "
                                  "(the collated mask is a list); only the causal-mask path exists")
    if args.report_to not in (None, "tensorboard"):
        warnings.warn(f"--report_to {args.report_to}: no tracker is attached here; metrics go to "
                      "<output_dir>/training.log and RUN_INFO")


def load_scheduler_config(path):
    """diffusers DDPMScheduler config (scheduler/scheduler_config.json) -> dict with the fields the step uses."""
    cfg = {"num_train_timesteps": 1000, "beta_start": 0.00085, "beta_end": 0.012,
           "beta_schedule": "scaled_linear", "prediction_type": "epsilon"}
    f = os.path.join(path, "scheduler", "scheduler_config.json")
    if os.path.exists(f):
        with open(f) as fh:
            raw = json.load(fh)
        cfg.update({k: raw[k] for k in cfg if k in raw})
    if cfg["beta_schedule"] != "scaled_linear":
        raise NotImplementedError(f"beta_schedule {cfg['beta_schedule']!r}: SD checkpoints use scaled_linear")
    return cfg


def load_tokenizer(args):
    """The reference's tokenizer load (train_textboost.py:630-638): --tokenizer_name (directory or hub id) or
    <checkpoint>/tokenizer.  The literal-id stand-in is handed out only for a synthetic checkpoint (marker file
    written by synthetic.write_pretrained) or under --synthetic_data; a real checkpoint whose tokenizer files are
    missing raises OSError."""
    from textboost_b200 import synthetic
    d = args.tokenizer_name or os.path.join(args.pretrained_model_name_or_path, "tokenizer")
    return synthetic.load_tokenizer(d, allow_literal=bool(getattr(args, "synthetic_data", False)))


def build_image_batches(args, tokenizer, rank, world):
    """The reference's train dataloader (train_textboost.py:856-890): PairedAugmentation -> TextBoostDataset ->
    Wrapper(drop_last=False).shuffle(seed).repeat() sharded by rank -> DataLoader(batch_size, collate_fn).  Returns an
    endless iterator of {"pixel_values" [B,3,S,S] fp32 in [-1,1] (or "sources" with --gpu_image_transforms),
    "input_ids" [B,L], "attention_mask"}; with
    --with_image_prior the class examples follow the instance ones in the same batch ([2B, ...])."""
    from textboost_b200.dataset import TextBoostDataset, Wrapper
    if args.augment in ("pda", "paug"):
        from textboost_b200.augment import PairedAugmentation
        augment_pipe = PairedAugmentation(hflip="inversion" if args.augment_inversion else "false",
                                          augment_prompt=args.augment_prompt, inversion=args.augment_inversion,
                                          p=args.augment_p, ops=args.augment_ops)
    elif args.augment == "custom_diff":
        raise NotImplementedError("--augment custom_diff: the reference imports a CustomDiffAugment class that its "
                                  "own textboost.augment package does not define (train_textboost.py:867)")
    else:
        augment_pipe = None
    dataset = TextBoostDataset(concepts_list=args.concepts_list, tokenizer=tokenizer, num_instance=args.num_samples,
                               template=args.template,
                               prior_data_root=args.class_data_dir if args.with_image_prior else None,
                               class_token=args.class_token,
                               num_prior=args.num_prior_images, size=args.resolution, center_crop=args.center_crop,
                               augment_pipe=augment_pipe, device_transforms=args.gpu_image_transforms,
                               cache_decoded=args.gpu_image_transforms, device_augment=args.gpu_augment)
    if len(dataset) == 0:
        raise ValueError("no instance images found")
    stream = Wrapper(dataset, drop_last=False, rank=rank, world_size=world).shuffle(seed=args.seed).repeat()
    # with the augmentation deferred to the GPU an item costs ~0.2 ms of host time: worker processes would only add a
    # fork and a shared-memory copy of each decoded base image per item
    workers = 0 if args.gpu_augment else args.dataloader_num_workers
    loader = torch.utils.data.DataLoader(stream, batch_size=args.train_batch_size,
                                         collate_fn=lambda ex: TextBoostDataset.collate_fn(ex, args.with_image_prior),
                                         num_workers=workers)
    RUN_INFO["dataloader_workers"] = workers
    RUN_INFO["instance_images"] = len(dataset)
    return iter(loader)


def log_validation(text_encoder, tokenizer, unet, vae, args, device, global_step):
    """train_textboost.py:453-531: sample ``num_validation_images`` images per validation prompt with the modules being
    trained (25 DPM-Solver++ steps, guidance 7.5); ``<i>`` in a prompt stands for concept i's placeholder tokens."""
    from textboost_b200 import pipeline as _pl
    from textboost_b200.pipeline import DiffusionPipeline
    logger.info(f"Running validation... \n Generating {args.num_validation_images} images with prompt:"
                f" {args.validation_prompts}.")
    pipeline = DiffusionPipeline.from_pretrained(args.pretrained_model_name_or_path, vae=vae, tokenizer=tokenizer,
                                                 text_encoder=text_encoder, unet=unet, safety_checker=None,
                                                 revision=args.revision, variant=args.variant)
    # train_textboost.py:483-495: a learned variance is mapped onto "fixed_small", then --validation_scheduler picks
    # the class (DPMSolverMultistepScheduler by default, DDPMScheduler)
    scheduler_args = {}
    cfg = pipeline.scheduler.config
    variance_type = cfg.get("variance_type") if isinstance(cfg, dict) else getattr(cfg, "variance_type", None)
    if variance_type is not None:
        scheduler_args["variance_type"] = "fixed_small" if variance_type in ("learned", "learned_range") else variance_type
    pipeline.scheduler = getattr(_pl, args.validation_scheduler).from_config(cfg, **scheduler_args)
    pipeline.set_progress_bar_config(disable=True)
    generator = None if args.seed is None else torch.Generator(device=device).manual_seed(args.seed)
    images = []
    for validation_prompt in args.validation_prompts:
        for i, placeholder in enumerate(args.placeholder_token):
            validation_prompt = validation_prompt.replace(f"<{i}>", " ".join(placeholder))
        images.extend(pipeline(prompt=validation_prompt, num_images_per_prompt=args.num_validation_images,
                               num_inference_steps=25, generator=generator).images)
    RUN_INFO.setdefault("validation_steps", []).append(global_step)
    return images


class ScalarTracker:
    """accelerate's tracker slice the reference uses (train_textboost.py:556-567, 944-945, 1018-1019, 1150, 1232, 1270):
    ``log({name: value}, step)`` and validation images, written as TensorBoard event files under
    ``<output_dir>/<logging_dir>/textboost`` when ``--report_to tensorboard`` (the default) and the package imports;
    otherwise a no-op (the training log and RUN_INFO carry the same numbers)."""

    def __init__(self, args, enabled: bool):
        self.writer = None
        if enabled and args.report_to == "tensorboard":
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(log_dir=os.path.join(args.output_dir, args.logging_dir, "textboost"))
            except Exception as e:  # a missing / broken tensorboard must never stop a training run
                warnings.warn(f"tensorboard tracker unavailable ({e}); metrics stay in training.log")

    def log(self, values: dict, step: int):
        if self.writer is not None:
            for name, value in values.items():
                self.writer.add_scalar(name, float(value), global_step=step)

    def images(self, tag: str, images, step: int):
        if self.writer is not None and images:
            import numpy as np
            self.writer.add_images(tag, np.stack([np.asarray(img) for img in images]), step, dataformats="NHWC")

    def close(self):
        if self.writer is not None:
            self.writer.close()
            self.writer = None


def save_learned_embeddings(text_encoder, added_tokens, aug_token_dict, directory):
    """train_textboost.py:1188-1209 / :1245-1266: one ``{token}.bin`` per added token (placeholder rows saved
    as [D], augmentation rows as [1, D])."""
    os.makedirs(directory, exist_ok=True)
    weight = text_encoder.get_input_embeddings().weight
    for token, token_id in added_tokens.items():
        name = token.replace("<", "").replace(">", "")
        torch.save({token: weight[token_id].detach().cpu()}, os.path.join(directory, f"{name}.bin"))
    for token, token_id in (aug_token_dict or {}).items():
        name = token.replace("<", "").replace(">", "")
        torch.save({token: weight[token_id:token_id + 1].detach().cpu()}, os.path.join(directory, f"{name}.bin"))


def save_checkpoint(trainer, text_encoder, step, directory, gen_state):
    os.makedirs(directory, exist_ok=True)
    st = {"step": step, "params": trainer.te.state.params.detach().cpu(),
          "optimizer": {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in trainer.opt.state_dict().items()},
          "generator": gen_state}
    if trainer.opt_unet is not None:  # --unet_params_to_train crossattn_kv: the third parameter group
        st["unet_lora_params"] = trainer.opt_unet.params.detach().cpu()
        st["unet_optimizer"] = {k: (v.cpu() if torch.is_tensor(v) else v)
                                for k, v in trainer.opt_unet.state_dict().items()}
        save_unet_lora(trainer.unet.kv_lora, directory)  # the hook's LoraLoaderMixin.save_lora_weights (:753-778)
    torch.save(st, os.path.join(directory, "state.pt"))
    text_encoder.save_pretrained(os.path.join(directory, "text_encoder"))


def save_unet_lora(kv_lora, directory):
    """The UNet adapter as ``pytorch_lora_weights.safetensors`` with the ``unet.`` key prefix of
    LoraLoaderMixin.save_lora_weights (train_textboost.py:753-778): "unet.<module>.lora_A.weight" / "lora_B.weight"."""
    from safetensors.torch import save_file
    os.makedirs(directory, exist_ok=True)
    save_file({"unet." + k: v.cpu().contiguous() for k, v in kv_lora.state_dict().items()},
              os.path.join(directory, "pytorch_lora_weights.safetensors"))


def load_checkpoint(trainer, directory, device):
    # state.pt holds tensors, ints and floats only (save_checkpoint): nothing is unpickled from the output directory
    st = torch.load(os.path.join(directory, "state.pt"), map_location="cpu", weights_only=True)
    trainer.te.state.params.copy_(st["params"].to(device))
    trainer.opt.load_state_dict({k: (v.to(device) if torch.is_tensor(v) else v) for k, v in st["optimizer"].items()})
    if trainer.opt_unet is not None:
        trainer.opt_unet.params.copy_(st["unet_lora_params"].to(device))
        trainer.opt_unet.load_state_dict({k: (v.to(device) if torch.is_tensor(v) else v)
                                          for k, v in st["unet_optimizer"].items()})
    return st["step"], st["generator"]


def _rotate_checkpoints(output_dir, limit):
    if limit is None:
        return
    ck = sorted((d for d in os.listdir(output_dir) if d.startswith("checkpoint")), key=lambda x: int(x.split("-")[1]))
    if len(ck) >= limit:
        for d in ck[:len(ck) - limit + 1]:
            shutil.rmtree(os.path.join(output_dir, d))


# ------------------------------------------------------------------------------------ main
def main(args):
    from textboost_b200 import dp, synthetic
    from textboost_b200.lora import LoraConfig
    from textboost_b200.text_encoder import TextBoostModel
    from textboost_b200.trainer import TextBoostTrainer, timestep_probs
    from textboost_b200.unet_model import UNet2DConditionModel
    from textboost_b200.utils import add_augmentation_tokens, add_token
    import copy

    _unsupported(args)
    # weight_dtype (train_textboost.py:928-933): selects the fp16 or the bf16 build of the library for this process
    from textboost_b200 import precision
    precision.set_policy(args.mixed_precision or "fp16")
    weight_dtype = precision.POLICY.act
    if not torch.cuda.is_available():
        raise SystemExit("train_textboost.py needs a B200: the CUDA library is the product, there is no CPU path")
    rank, world, local_rank = dp.init_from_env()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    is_main = rank == 0
    os.makedirs(args.output_dir, exist_ok=True)
    logging.basicConfig(level=logging.INFO if is_main else logging.WARNING,
                        format="%(asctime)s - %(levelname)s - %(name)s - %(message)s",
                        handlers=[logging.StreamHandler()] + (
                            [logging.FileHandler(os.path.join(args.output_dir, "training.log"))] if is_main else []))
    if args.seed is not None:  # same seed on every rank (train_textboost.py:598-601, SURVEY.md trap 10)
        import random
        import numpy as np
        random.seed(args.seed)  # accelerate.set_seed: python, numpy and torch streams (the augmentation draws from
        np.random.seed(args.seed)  # the first two, the random crop from the third)
        torch.manual_seed(args.seed)

    if args.concepts_list is None:  # train_textboost.py:602-615
        args.concepts_list = [{"instance_token": args.instance_token, "class_token": args.class_token,
                               "instance_data_dir": args.instance_data_dir, "class_data_dir": args.class_data_dir,
                               "placeholder_token": args.placeholder_token, "initializer_token": args.initializer_token}]
    else:
        with open(args.concepts_list) as f:
            args.concepts_list = json.load(f)

    path = args.pretrained_model_name_or_path
    tokenizer = load_tokenizer(args)
    sched = load_scheduler_config(path)
    text_encoder = TextBoostModel.from_pretrained(path, subfolder="text_encoder", revision=args.revision,
                                                  variant=args.variant)
    unet = UNet2DConditionModel.from_pretrained(path, subfolder="unet", revision=args.revision, variant=args.variant)

    # null embedding (train_textboost.py:649 + SURVEY.md D6)
    D = text_encoder.config.hidden_size
    asset = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "null_emb_sd21base.pt")
    null = None
    if args.null_embedding is not None:
        null = torch.load(args.null_embedding, map_location="cpu", weights_only=True)
    elif os.path.exists(asset):
        cand = torch.load(asset, map_location="cpu", weights_only=True)
        null = cand if cand.shape[-1] == D else None
    if null is None:  # derive it: the frozen encoder's hidden states for the empty prompt
        probe = copy.deepcopy(text_encoder).to(device)
        ids = torch.full((1, text_encoder.config.max_position_embeddings), 49407, dtype=torch.int64)
        ids[0, 0] = 49406
        with torch.no_grad():
            null = probe(ids.to(device))[0][0].float().cpu()
        del probe
    text_encoder.set_null_embedding(null)
    original_text_encoder = copy.deepcopy(text_encoder).eval().requires_grad_(False)

    # placeholder / augmentation tokens (train_textboost.py:659-694)
    added_tokens, placeholder_token_ids = {}, []
    for concept in args.concepts_list:
        toks, ids = add_token(text_encoder, tokenizer, concept["placeholder_token"], concept["initializer_token"])
        placeholder_token_ids += ids
        added_tokens.update(dict(zip(toks, ids)))
        concept["instance_token"] = concept["placeholder_token"] = toks
    args.placeholder_token = [c["placeholder_token"] for c in args.concepts_list]  # train_textboost.py:662-676
    aug_token_dict = {}
    if args.augment_inversion:
        _, aug_token_dict = add_augmentation_tokens(text_encoder, tokenizer,
                                                    aug_type="style" if args.augment_ops == "style" else "object")

    unet.eval().requires_grad_(False)
    text_encoder.requires_grad_(False)
    if args.lora_rank > 0:  # train_textboost.py:700-709
        text_encoder.text_model.encoder.requires_grad_(True)
        text_encoder.add_adapter(LoraConfig(r=args.lora_rank, lora_alpha=args.lora_rank, init_lora_weights="gaussian",
                                            target_modules=["q_proj", "k_proj", "v_proj"]))
    # --lora_rank 0: no adapter and the encoder stays frozen (the reference only un-freezes it inside the branch
    # above, :700-701), so the added embedding rows are the only trainable state -- plain textual inversion; the
    # optimiser's second parameter group is empty and the gradient clipping (:1128-1133) has nothing to clip
    text_encoder.get_input_embeddings().requires_grad_(True)
    n_lora = sum(p.numel() for n, p in text_encoder.named_parameters() if "lora" in n)
    logger.info(f"trainable: {n_lora} LoRA floats + {len(added_tokens) + len(aug_token_dict)} embedding rows")

    if args.scale_lr:  # train_textboost.py:818-821
        args.learning_rate *= args.gradient_accumulation_steps * args.train_batch_size * world

    # mean row norm of the resized embedding matrix, before training (train_textboost.py:1003-1021)
    mean_norm = text_encoder.get_input_embeddings().weight.norm(dim=-1).mean().item()

    text_encoder.to(device)
    unet.to(device, dtype=weight_dtype)
    original_text_encoder.to(device, dtype=weight_dtype)
    unet_lora = None
    if args.unet_params_to_train == "crossattn_kv" and args.lora_rank > 0:  # train_textboost.py:712-721
        unet_lora = unet.engine.add_cross_kv_lora(args.lora_rank, alpha=args.lora_rank,
                                                  seed=(args.seed if args.seed is not None else 0) + 17)
        logger.info(f"Added LoRA to U-Net: {unet_lora.n_adapters} adapters, {unet_lora.params.numel()} floats")
    trainer = TextBoostTrainer(
        unet.engine, text_encoder.engine, original_text_encoder.engine if args.kpl_weight > 0 else None,
        learning_rate=args.learning_rate, emb_learning_rate=args.emb_learning_rate, adam_beta1=args.adam_beta1,
        adam_beta2=args.adam_beta2, adam_weight_decay=args.adam_weight_decay, adam_epsilon=args.adam_epsilon,
        max_grad_norm=args.max_grad_norm, kpl_weight=args.kpl_weight, kpl_type=args.kpl_type,
        prediction_type=sched["prediction_type"],
        mixing=("style" if args.augment_ops == "style" else "object") if args.mixing else None,
        mean_norm=mean_norm, mixed_precision=args.mixed_precision or "fp16",
        image_prior_weight=args.image_ppl_weight if args.with_image_prior else None,
        lr_scheduler=args.lr_scheduler, lr_warmup_steps=args.lr_warmup_steps, max_train_steps=args.max_train_steps,
        gradient_accumulation_steps=args.gradient_accumulation_steps,
        num_train_timesteps=sched["num_train_timesteps"], beta_start=sched.get("beta_start", 0.00085),
        beta_end=sched.get("beta_end", 0.012))

    # ---- data: the image front end (dataset -> augmentation -> VAE encoder), a latents file, or synthetic latents
    B = args.train_batch_size
    latent = args.resolution // 8
    vae = image_batches = None
    if not args.latents_file and not args.synthetic_data:
        if not all(c.get("instance_data_dir") for c in args.concepts_list):
            raise ValueError("no training data: pass --instance_data_dir / --concepts_list (image front end), "
                             "--latents_file or --synthetic_data")
        from textboost_b200.vae import AutoencoderKL
        vae = AutoencoderKL.from_pretrained(path, subfolder="vae", revision=args.revision, variant=args.variant)
        vae.eval().requires_grad_(False)
        vae.to(device, dtype=torch.float32)
        image_batches = build_image_batches(args, tokenizer, rank, world)
        lat_all = ids_all = prior_all = None
        logger.info(f"image front end: {RUN_INFO['instance_images']} instance images, augment={args.augment}")
    elif args.latents_file:
        data = torch.load(args.latents_file, map_location="cpu", weights_only=True)
        lat_all, ids_all = data["latents"].float(), data["input_ids"].long()
        prior_all = data.get("prior_ids")
    elif args.synthetic_data:
        n = max(B * world, 8)
        g = torch.Generator().manual_seed(args.seed)
        lat_all = torch.randn(n, 4, latent, latent, generator=g)
        ids_all = synthetic.instance_ids(n, placeholder_token_ids[0], text_encoder.config.max_position_embeddings)
        prior_all = synthetic.prior_ids(max(args.num_prior_images, B * world), args.seed + 1,
                                        text_encoder.config.max_position_embeddings, args.null_prob)
    prior_stream = None
    if args.kpl_weight > 0 and os.path.exists(args.prior_prompts_file):
        # the reference's prior-prompt pipeline (train_textboost.py:893-909): human-written prompts, null / template
        # mixing, shuffled + repeated + sharded by rank; batch_size prompts per step
        from textboost_b200 import prompts as P
        import itertools
        src = P.HumanPromptSource(tokenizer, args.prior_prompts_file, num_samples=None)
        prior_ds = P.PriorPrompts(src, tokenizer, additional_template=args.template,
                                  additional_category=args.class_token, null_prob=args.null_prob)
        stream = iter(P.ShardedStream(prior_ds, drop_last=True, rank=rank, world_size=world)
                      .shuffle(seed=args.seed).repeat())
        prior_stream = (P.PriorPrompts.collate_fn(list(itertools.islice(stream, args.train_batch_size)))["input_ids"]
                        for _ in itertools.count())
        logger.info(f"prior prompts: {len(src)} from {args.prior_prompts_file}")
        RUN_INFO["prior_prompts"] = len(src)
    if args.kpl_weight > 0 and prior_all is None and prior_stream is None:
        raise ValueError("kpl_weight > 0 needs prior prompts: --prior_prompts_file (JSONL) or 'prior_ids' in "
                         "--latents_file")
    if image_batches is None:
        lat_all, ids_all = lat_all.to(device), ids_all.to(device)
    prior_all = prior_all.to(device) if prior_all is not None else None
    gen = torch.Generator(device=device)
    gen.manual_seed(args.seed)
    T = sched["num_train_timesteps"]
    p_t = None if args.disable_weighted_sample else timestep_probs(trainer.acp)

    step = 0
    if args.resume_from_checkpoint:  # train_textboost.py:960-981, made to work
        ck = args.resume_from_checkpoint
        if ck == "latest":
            dirs = sorted((d for d in os.listdir(args.output_dir) if d.startswith("checkpoint")),
                          key=lambda x: int(x.split("-")[1]))
            ck = os.path.join(args.output_dir, dirs[-1]) if dirs else None
        if ck is None or not os.path.isdir(ck):
            logger.info(f"Checkpoint '{args.resume_from_checkpoint}' does not exist. Starting a new training run.")
        else:
            step, gstate = load_checkpoint(trainer, ck, device)
            gen.set_state(gstate)
            logger.info(f"Resuming from checkpoint {ck} at step {step}")

    def draw(step_idx):
        """rank r takes rows [r*B, (r+1)*B) of the step's global batch (dataset.py:846-870 sharding)."""
        if image_batches is not None:  # train_textboost.py:1027-1037: pixels -> VAE posterior sample * scaling_factor
            batch = next(image_batches)
            if "sources" in batch:  # --gpu_image_transforms: resize / crop / normalise on the GPU, byte-exact
                from textboost_b200.image_ops import batch_to_pixel_values
                pixels = batch_to_pixel_values(batch["sources"], device)
            else:
                pixels = batch["pixel_values"].to(device, non_blocking=True)
            lat = vae.engine.encode_latents(pixels, generator=gen)
            ids = batch["input_ids"].to(device, non_blocking=True)
        else:
            n = lat_all.shape[0]
            idx = (torch.arange(B * world, device=device) + step_idx * B * world) % n
            idx = idx[rank * B:(rank + 1) * B]
            lat, ids = lat_all[idx], ids_all[idx]
        noise = torch.randn(lat.shape, generator=gen, device=device)
        bsz = lat.shape[0]  # 2B with --with_image_prior (train_textboost.py:1045)
        if p_t is None:
            t = torch.randint(0, T, (bsz,), generator=gen, device=device)
        else:
            t = torch.multinomial(p_t, bsz, replacement=True, generator=gen)
        pri = None
        if prior_stream is not None:
            pri = next(prior_stream).to(device, non_blocking=True)
        elif prior_all is not None and args.kpl_weight > 0:
            pidx = (torch.arange(B * world, device=device) + step_idx * B * world) % prior_all.shape[0]
            pri = prior_all[pidx[rank * B:(rank + 1) * B]]
        return lat, noise, t, ids, pri

    tracker = ScalarTracker(args, is_main)
    tracker.log({"mean_norm": mean_norm}, 0)
    logger.info("***** Running training *****")
    logger.info(f"  Instantaneous batch size per device = {B}")
    logger.info(f"  Total train batch size (w. parallel) = {B * world}")
    logger.info(f"  Total optimization steps = {args.max_train_steps}")
    logger.info(f"  Mean norm: {mean_norm}")
    run = trainer.step  # first step eager (configures kernel attributes), then the captured CUDA graph
    start = time.perf_counter()
    loss_val = float("nan")
    first = step
    accum = args.gradient_accumulation_steps  # accelerator.accumulate (train_textboost.py:1039): one optimiser
    micro = step * accum                      # step per `accum` dataloader batches
    while step < args.max_train_steps:
        for k in range(accum):
            batch = draw(micro)
            if step == first + 1 and k == 0 and args.max_train_steps - step > 2:
                # capture over static copies of THIS batch, then replay it: no batch is drawn and dropped
                run = trainer.capture(*batch, warmup=0)
            loss = run(*batch, sync_gradients=(k == accum - 1))
            micro += 1
        step += 1
        if step % args.log_every == 0 or step == args.max_train_steps:
            loss_val = loss.item()  # the reference syncs every step (:1230); here every --log_every steps
            added_norm = trainer.added_norm.item()
            lr_now = trainer.opt.get_last_lr()[0]  # lr_scheduler.get_last_lr()[0] of the reference (:1230): group 0 = embeddings
            logger.info(f"step {step} loss {loss_val:.6f} lr {lr_now:.3e} added_embedding_norm {added_norm:.4f}")
            tracker.log({"loss": loss_val, "lr": lr_now, "added_embedding_norm": added_norm}, step)
        if is_main and args.validation_prompts and step % args.validation_steps == 0:  # train_textboost.py:1213-1228
            if vae is None or vae.decoder_engine is None:
                from textboost_b200.vae import AutoencoderKL
                vae = AutoencoderKL.from_pretrained(path, subfolder="vae", revision=args.revision,
                                                    variant=args.variant).to(device, dtype=torch.float32)
            images = log_validation(text_encoder, tokenizer, unet, vae, args, device, step)
            tracker.images("validation", images, step)
            if images:
                from inference import make_image_grid
                grid = make_image_grid(images, len(args.validation_prompts), args.num_validation_images)
                grid.save(os.path.join(args.output_dir, f"validation_{step}.jpg"))
        if is_main and step % args.checkpointing_steps == 0:
            _rotate_checkpoints(args.output_dir, args.checkpoints_total_limit)
            ck = os.path.join(args.output_dir, f"checkpoint-{step}")
            save_checkpoint(trainer, text_encoder, step, ck, gen.get_state())
            save_learned_embeddings(text_encoder, added_tokens, aug_token_dict, ck)
            logger.info(f"Saved state to {ck}")
    torch.cuda.synchronize()
    if is_main:
        if args.lora_rank > 0:
            text_encoder.to(torch.float32).save_pretrained(os.path.join(args.output_dir, "text_encoder"),
                                                           safe_serialization=not args.no_safe_serialization)
        save_learned_embeddings(text_encoder, added_tokens, aug_token_dict, args.output_dir)
        if unet_lora is not None:
            # the reference writes the whole peft-wrapped UNet (unet.save_pretrained, :1237-1239); the frozen base is
            # the checkpoint the run started from, so only the adapter is written here
            save_unet_lora(unet_lora, os.path.join(args.output_dir, "unet"))
    tracker.close()
    _st = trainer.opt_state.tolist()
    RUN_INFO.update(precision=precision.POLICY.name, loss_scale=_st[0], skipped_steps=int(_st[8]))
    logger.info(f"Training took {time.perf_counter() - start:.2f} seconds")
    if world > 1:
        # the captured graph holds the NCCL communicator: release it before the ranks part (destroying the process
        # group with the graph alive blocks), and leave NCCL teardown to process exit
        trainer._graph = None
        run = None  # noqa: F841
        torch.cuda.synchronize()
        torch.distributed.barrier()
        logging.shutdown()
        os._exit(0)
    return loss_val


if __name__ == "__main__":
    main(parse_args())
