"""One TextBoost training step, fused end to end on the B200 kernels.

Restates the body of the reference loop (/root/reference/train_textboost.py:1041-1149) as a fixed kernel
sequence with no autograd graph and no host synchronisation:

  add_noise -> text encoder (instance + prior prompts in one pass, LoRA fused) -> UNet forward ->
  MSE (+grad) -> UNet dgrad to encoder_hidden_states -> frozen text encoder on the prior prompts ->
  knowledge-preservation loss (+grad) -> text-encoder backward (LoRA A/B + added embedding rows) ->
  ONE all-reduce of the flat gradient buffer -> fused unscale / clip / AdamW / renorm / GradScaler.

Differences from the reference that do not change results: the DDP all-reduce covers
only LoRA + added rows (the reference reduces the whole embedding matrix and then zeroes the frozen rows,
D5); the weight decay of the frozen embedding rows is a lazily applied scalar (D8); instance and prior
prompts share one encoder pass (the ops are row-independent).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _cabi as C
from . import ops
from .clip import ClipEngine
from .dp import GradSync
from .optim import FlatAdamW, FusedAdamW
from .unet import UNetEngine

from .precision import POLICY
F32 = torch.float32


def alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, device="cpu"):
    """SD scheduler config (scaled_linear): DDPMScheduler as used at train_textboost.py:644."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=F32) ** 2
    return torch.cumprod(1.0 - betas, dim=0).to(device)


def timestep_probs(acp: torch.Tensor) -> torch.Tensor:
    """train_textboost.py:991-997 (reachable here; the reference's flag keeps it dead code, D3)."""
    logsnr = (acp / (1 - acp)).log()
    w = -logsnr + logsnr.max()
    return w / w.sum()


class TextBoostTrainer:
    def __init__(self, unet: UNetEngine, text_encoder: ClipEngine,
                 original_text_encoder: Optional[ClipEngine] = None, *, learning_rate=5e-5,
                 emb_learning_rate=1e-3, adam_beta1=0.9, adam_beta2=0.999, adam_weight_decay=1e-2,
                 adam_epsilon=1e-8, max_grad_norm=1.0, kpl_weight=0.1, kpl_type="cos",
                 prediction_type="epsilon", mixing=None, mean_norm: Optional[float] = None,
                 mixed_precision: Optional[str] = None, process_group=None, image_prior_weight: Optional[float] = None,
                 lr_scheduler="constant", lr_warmup_steps=0, max_train_steps=0, gradient_accumulation_steps=1,
                 num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
        self.unet, self.te, self.te0 = unet, text_encoder, original_text_encoder
        # the engines were built for the process's precision policy (precision.py); the GradScaler follows it
        mixed_precision = mixed_precision or POLICY.name
        if mixed_precision != POLICY.name:
            raise ValueError(f"mixed_precision={mixed_precision!r} but the process's precision policy is {POLICY.name!r} "
                             "(textboost_b200.precision.set_policy before building the engines)")
        # --with_image_prior (train_textboost.py:1077-1094): the batch is [instance | class] halves and the loss is
        # mse(instance half) + image_prior_weight * mse(class half); None = the plain single-part loss
        self.image_prior_weight = image_prior_weight
        self.dev = text_encoder.device
        self.kpl_weight, self.kpl_kind = kpl_weight, {"cos": 0, "mse": 1}[kpl_type]
        self.v_pred = {"epsilon": False, "v_prediction": True}[prediction_type]
        assert kpl_weight <= 0 or original_text_encoder is not None
        self.sync = GradSync(process_group)
        self.world = self.sync.world
        if gradient_accumulation_steps > 1 and self.world > 1:
            # train_textboost.py:573-577
            raise ValueError("Gradient accumulation is not supported when training the text encoder in distributed "
                             "training. Please set gradient_accumulation_steps to 1.")
        self.opt = FusedAdamW(text_encoder, lr=learning_rate, emb_lr=emb_learning_rate,
                              betas=(adam_beta1, adam_beta2), weight_decay=adam_weight_decay, eps=adam_epsilon,
                              max_grad_norm=max_grad_norm, mean_norm=mean_norm, mixing=mixing,
                              mixed_precision=mixed_precision, world_size=self.world, lr_scheduler=lr_scheduler,
                              lr_warmup_steps=lr_warmup_steps, max_train_steps=max_train_steps,
                              gradient_accumulation_steps=gradient_accumulation_steps)
        # --unet_params_to_train crossattn_kv (train_textboost.py:712-721, 838-841): the UNet engine carries a K/V
        # adapter (UNetEngine.add_cross_kv_lora); its flat buffer is the third parameter group: the LoRA learning
        # rate, weight decay, no clipping (:1128-1133 clip the text encoder only).  In the reference the mode only runs
        # without a GradScaler (its fp16 policy raises in GradScaler.unscale_ on the fp16 adapter tensors): here it is
        # tied to the bf16 policy, whose loss scale is the constant 1 both optimiser calls assume.
        self.opt_unet = None
        if getattr(unet, "kv_lora", None) is not None:
            if mixed_precision != "bf16":
                raise NotImplementedError("the UNet K/V adapter (--unet_params_to_train crossattn_kv) needs "
                                          "--mixed_precision bf16: the reference's fp16 path fails in "
                                          "GradScaler.unscale_ on the fp16 adapter tensors, and fp32 is not built")
            self.opt_unet = FlatAdamW(unet.kv_lora.params, unet.kv_lora.grads, lr=learning_rate,
                                      betas=(adam_beta1, adam_beta2), weight_decay=adam_weight_decay, eps=adam_epsilon,
                                      world_size=self.world, lr_scheduler=lr_scheduler,
                                      lr_warmup_steps=lr_warmup_steps, max_train_steps=max_train_steps,
                                      gradient_accumulation_steps=gradient_accumulation_steps)
        # the checkpoint's scheduler/scheduler_config.json (DDPMScheduler.from_pretrained, train_textboost.py:644)
        self.num_train_timesteps = int(num_train_timesteps)
        self.acp = alphas_cumprod(self.num_train_timesteps, beta_start, beta_end, device=self.dev)
        self.loss = torch.zeros(1, device=self.dev, dtype=F32)
        self._graph = None
        import os
        self.separate_encoder_passes = bool(os.environ.get("TB_SEPARATE_ENCODER_PASSES"))  # A/B switch

    # optimiser state under the names the tests / bench use
    lr = property(lambda self: self.opt.param_groups[1]["lr"])
    emb_lr = property(lambda self: self.opt.param_groups[0]["lr"])
    b1 = property(lambda self: self.opt.betas[0])
    b2 = property(lambda self: self.opt.betas[1])
    wd = property(lambda self: self.opt.weight_decay)
    eps = property(lambda self: self.opt.eps)
    max_grad_norm = property(lambda self: self.opt.max_grad_norm)
    mixing = property(lambda self: self.opt.mixing)
    mean_norm = property(lambda self: self.opt.mean_norm)
    opt_state = property(lambda self: self.opt.state)
    added_norm = property(lambda self: self.opt.added_norm)
    exp_avg = property(lambda self: self.opt.exp_avg)
    exp_avg_sq = property(lambda self: self.opt.exp_avg_sq)

    # ------------------------------------------------------------------ pieces (also used by tests)
    def forward_backward(self, latents, noise, timesteps, input_ids, prior_ids=None):
        """Everything up to (not including) the all-reduce: fills te.state.grads, self.loss.

        The knowledge-preservation branch (trainable + frozen encoder on the prior prompts, the KPL term and
        its backward, train_textboost.py:1096-1106) does not depend on the UNet, so it is enqueued on a side
        stream and overlaps the UNet forward / activation-backward; the two branches meet before the
        instance-prompt encoder backward.  Inside capture() the fork / join become graph edges."""
        te, unet = self.te, self.unet
        B = latents.shape[0]
        use_kpl = self.kpl_weight > 0 and prior_ids is not None
        scale = self.opt_state[0:1]
        main = torch.cuda.current_stream()
        te.pack_lora()
        ops.zero_(self.loss)
        noisy, target = ops.add_noise(latents, noise, timesteps, self.acp, self.v_pred)
        # The text-encoder passes run on a side stream: first the instance prompts (the UNet needs them at its first
        # cross-attention: an event, not a join), then the whole knowledge-preservation branch.  The main stream goes
        # straight into the UNet, whose prefix (conv_in, time MLP, first resnet, first self-attention) does not read
        # the text: the 616-row encoder kernels fill a fraction of the SMs, the UNet prefix takes the rest.
        side = self._side_stream()
        first = self._first_stream()  # the instance prompts' forward: the UNet waits for it -> high priority
        ops.zero_(self._loss_kpl)
        first.wait_stream(main)
        side.wait_stream(main)
        # instance and prior prompts go through the trainable encoder in ONE pass when they have the same length: the
        # encoder's kernels are latency-bound at 616 rows and cost the same at 1232, so the second pass (87 launches
        # competing with the UNet for SMs) disappears; the two halves' backwards still run separately
        one_pass = use_kpl and prior_ids.shape[1:] == input_ids.shape[1:] and not self.separate_encoder_passes
        with torch.cuda.stream(first):
            if one_pass:
                ids_all = torch.empty((B + prior_ids.shape[0],) + tuple(input_ids.shape[1:]), device=input_ids.device,
                                      dtype=input_ids.dtype)
                ids_all[:B].copy_(input_ids)  # two device-to-device copies (memcpy nodes), not a cat kernel
                ids_all[B:].copy_(prior_ids)
                h_all = te.forward(ids_all, save_for_backward=True)
                ctx_i, ctx_p = te.split_ctx(te.pop_ctx(), B)
                h, hp = h_all[:B], h_all[B:]
            else:
                h = te.forward(input_ids, save_for_backward=True)  # fp32 [B, L, D]
                ctx_i = te.pop_ctx()
            _, L, D = h.shape
            ehs = ops.cast_f32_f16(h.reshape(B * L, D)).view(B, L, D)
            if ehs.is_cuda:
                ehs.record_stream(main)
            ehs_ready = torch.cuda.Event()
            ehs_ready.record(first)
        with torch.cuda.stream(side):
            if use_kpl:
                if one_pass:
                    side.wait_event(ehs_ready)
                else:
                    hp = te.forward(prior_ids, save_for_backward=True)
                    ctx_p = te.pop_ctx()
                h0 = self.te0.forward(prior_ids)
                Bp, L, D = hp.shape
                d_hp = ops.zeros((Bp, L, D), self.dev)
                C.call("tb_kpl_fwd_bwd", C.ptr(hp), C.ptr(h0), Bp * L, D, self.kpl_kind, float(self.kpl_weight),
                       C.ptr(scale), C.ptr(self._loss_kpl), C.ptr(d_hp), C.stream_ptr())
                te.backward(d_hp, ctx=ctx_p)
        _, L, D = h.shape
        pred = unet.forward(noisy, timesteps, ehs, ehs_ready=ehs_ready)
        if self.image_prior_weight is None:
            dpred = ops.mse_fwd_bwd(pred, target, self.loss, 1.0, scale)
        else:
            if B % 2:
                raise ValueError("with an image prior the batch is [instance | class] halves: even batch size")
            half = B // 2
            dpred = torch.empty_like(pred)
            ops.mse_fwd_bwd(pred[:half], target[:half], self.loss, 1.0, scale, out=dpred[:half])
            ops.mse_fwd_bwd(pred[half:], target[half:], self.loss, float(self.image_prior_weight), scale,
                            out=dpred[half:])
        d_h = ops.zeros((B, L, D), self.dev)
        unet.backward(dpred, d_h)
        main.wait_stream(side)  # gradient accumulation into state.grads is serialised from here on
        if first is not side:
            main.wait_stream(first)
        if use_kpl:
            ops.axpy_(self.loss, self._loss_kpl)
        te.backward(d_h, ctx=ctx_i)
        self._pred = pred
        return self.loss

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.dev)
            self._loss_kpl = torch.zeros(1, device=self.dev, dtype=F32)
        return self._side

    def _first_stream(self):
        """Stream of the instance-prompt encoder forward.  The UNet's first cross-attention waits for its result while
        the UNet prefix fills the machine with persistent kernels; a high-priority stream lets the 616-row encoder
        kernels take SMs at every kernel boundary of the prefix instead of queueing behind it.  The knowledge-
        preservation branch keeps its own normal-priority stream: nothing waits for it until the end of the UNet
        backward.  (TB_ONE_SIDE_STREAM=1: the single side stream of round 1, for comparison.)"""
        if getattr(self, "_first", None) is None:
            import os
            if os.environ.get("TB_ONE_SIDE_STREAM") or not torch.cuda.is_available():
                self._first = self._side_stream()
            else:
                self._first = torch.cuda.Stream(device=self.dev, priority=-1)
        return self._first

    def all_reduce(self):
        self.sync.all_reduce_(self.te.state.grads)
        if self.opt_unet is not None:
            self.sync.all_reduce_(self.opt_unet.grads)

    def optimizer_step(self):
        self.opt.step()
        if self.opt_unet is not None:
            self.opt_unet.step()

    # ------------------------------------------------------------------ the step
    def step(self, latents, noise, timesteps, input_ids, prior_ids=None, sync_gradients=True):
        """latents/noise fp32 [B,4,H,W]; timesteps int64 [B]; ids int64 [B,L].  Returns the device-side
        loss scalar (fp32[1]); nothing is synchronised.

        sync_gradients=False: a micro-batch of ``accelerator.accumulate`` (train_textboost.py:1039): its gradients are
        added to the flat buffer and the all-reduce / optimiser tail are left to the micro-batch that closes the
        window (the 1 / gradient_accumulation_steps of accelerate's backward is applied there)."""
        self.forward_backward(latents, noise, timesteps, input_ids, prior_ids)
        if sync_gradients:
            self.all_reduce()
            self.optimizer_step()
        return self.loss

    # ------------------------------------------------------------------ CUDA graph of the whole step
    def capture(self, latents, noise, timesteps, input_ids, prior_ids=None, warmup=2):
        """Capture step() into one CUDA graph over static input buffers; returns replay(inputs...)."""
        static = [t.clone() if t is not None else None
                  for t in (latents, noise, timesteps, input_ids, prior_ids)]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.step(*static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step(*static)
        self._graph = g
        g_acc = None
        if self.opt.gradient_accumulation_steps > 1:  # the micro-batches that only accumulate: a second graph
            g_acc = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_acc, pool=g.pool()):
                self.step(*static, sync_gradients=False)
            # (capturing executes nothing: the gradient buffer is untouched)

        # the optimiser's scalar hyper-parameters are arguments of the captured launch: a later edit of
        # param_groups[...]["lr"] (or betas / weight decay) would be silently ignored by the replay -- refuse instead
        baked = self.opt.hyperparameters()

        def replay(latents, noise, timesteps, input_ids, prior_ids=None, sync_gradients=True):
            if self.opt.hyperparameters() != baked:
                raise RuntimeError("optimiser hyper-parameters changed after capture(): they are baked into the CUDA "
                                   "graph; call capture() again (the --lr_scheduler runs on the device and needs no edit)")
            for dst, src in zip(static, (latents, noise, timesteps, input_ids, prior_ids)):
                if dst is not None and src is not None and dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
            (g if sync_gradients or g_acc is None else g_acc).replay()
            return self.loss

        self.static_inputs = static
        self._replay = replay
        return replay

    # ------------------------------------------------------------------ host-facing step
    def step_from_host(self, latents, noise, timesteps, input_ids, prior_ids=None) -> float:
        """One training step from HOST tensors (pinned memory makes the copies asynchronous): copies the
        batch to the device, runs the step (graph replay once capture() has been called) and reads the loss
        back, like the reference loop's ``loss.detach().item()`` (train_textboost.py:1230)."""
        host = (latents, noise, timesteps, input_ids, prior_ids)
        if self._graph is not None:
            for dst, src in zip(self.static_inputs, host):
                if dst is not None and src is not None:
                    dst.copy_(src, non_blocking=True)
            self._graph.replay()
        else:
            dev = [t.to(self.dev, non_blocking=True) if t is not None else None for t in host]
            self.step(*dev)
        if getattr(self, "_loss_host", None) is None:
            self._loss_host = torch.empty(1, dtype=F32, pin_memory=True)
        self._loss_host.copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self._loss_host[0])
