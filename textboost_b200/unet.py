"""Frozen SD UNet denoiser: forward and activation-backward (dgrad only) on the B200 kernels.

Drop-in for ``diffusers.UNet2DConditionModel`` on the TextBoost path
(/root/reference/train_textboost.py:654-656 load, :696 freeze, :937 fp16 cast, :1063-1067 call,
:1108 backward through it to ``encoder_hidden_states``).  The module graph follows diffusers 0.29.0
(SURVEY.md Appendix A.1); state-dict keys are the diffusers ones (Appendix A.4).

Design (B200-first, not a port):
  * activations are channels-last fp16 ([B,H,W,C] == [B,HW,C] tokens), so the Transformer2D reshape is free
    and every conv / linear is one tcgen05 GEMM (conv3x3 = implicit GEMM, TMA-gathered);
  * bias, time-embedding add, residual add and SiLU are GEMM epilogues; GroupNorm+SiLU is one HBM pass;
  * the UNet is frozen: backward is a hand-derived dgrad chain (no wgrad, no autograd graph), pruned
    above the first cross-attention, with d(encoder_hidden_states) accumulated in fp32 by the epilogue of
    the 16 cross-attention K/V dgrad GEMMs;
  * weights are stored twice (forward layout and pre-transposed / tap-flipped dgrad layout).
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Tuple

import torch

from . import _cabi as C
from . import ops

from .precision import POLICY
F32 = torch.float32


@dataclasses.dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    attention_head_dim: Tuple[int, ...] = (8, 8, 8, 8)  # heads per level (diffusers naming quirk)
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    sample_size: int = 64

    @staticmethod
    def sd15():
        return UNetConfig()

    @staticmethod
    def sd21():
        return UNetConfig(cross_attention_dim=1024, attention_head_dim=(5, 10, 20, 20),
                          use_linear_projection=True, sample_size=96)


def _conv_fwd_weight(w):  # [Cout,Cin,3,3] -> [Cout, 9*Cin] tap-major
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def _conv_dgrad_weight(w):  # -> [Cin, 9*Cout], taps flipped
    return w.flip(2, 3).permute(1, 2, 3, 0).reshape(w.shape[1], -1).contiguous()


class _Linear:
    """y = x W^T + b ; holds W [N,K] and W^T [K,N] (dgrad operand)."""

    def __init__(self, w, b=None):
        if w.dim() == 4:  # 1x1 conv
            w = w.reshape(w.shape[0], w.shape[1])
        self.w = w.contiguous()
        self.wt = w.t().contiguous()
        self.b = b

    def fwd(self, x, **kw):
        return ops.gemm(x, self.w, bias=self.b, **kw)

    def dgrad(self, dy, **kw):
        return ops.gemm(dy, self.wt, **kw)


class _Conv3:
    def __init__(self, w, b):
        self.wk = _conv_fwd_weight(w)
        self.wd = _conv_dgrad_weight(w)
        self.b = b

    def fwd(self, x, **kw):
        return ops.conv3x3(x, self.wk, bias=self.b, **kw)

    def dgrad(self, dy, **kw):
        return ops.conv3x3(dy, self.wd, **kw)


class _Resnet:
    arena = None  # ops.StatsArena of the running UNet pass (set by UNetEngine), None = per-call memset

    def __init__(self, sd, p, cfg: UNetConfig):
        self.G, self.eps = cfg.norm_num_groups, cfg.norm_eps
        self.n1 = (sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        self.c1 = _Conv3(sd[p + "conv1.weight"], sd[p + "conv1.bias"])
        self.t = _Linear(sd[p + "time_emb_proj.weight"], sd[p + "time_emb_proj.bias"])
        self.n2 = (sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        self.c2 = _Conv3(sd[p + "conv2.weight"], sd[p + "conv2.bias"])
        self.sc = None
        if p + "conv_shortcut.weight" in sd:
            self.sc = _Linear(sd[p + "conv_shortcut.weight"], sd[p + "conv_shortcut.bias"])

    def forward(self, x, temb_act, save, tp=None):
        """tp: this block's time_emb_proj(silu(temb)) [B, Cout], normally a column slice of the ONE GEMM that projects
        the time embedding for every ResnetBlock2D of the net (UNetEngine.forward)."""
        B, H, W, Cin = x.shape
        h, st1 = ops.groupnorm(x, *self.n1, self.G, self.eps, True, arena=self.arena)
        if tp is None:
            tp = self.t.fwd(temb_act)
        h2 = self.c1.fwd(h, rowvec=tp)
        h3, st2 = ops.groupnorm(h2, *self.n2, self.G, self.eps, True, arena=self.arena)
        if self.sc is not None:
            res = self.sc.fwd(x.view(-1, Cin)).view(B, H, W, -1)
        else:
            res = x
        out = self.c2.fwd(h3, residual=res)
        if save:
            self.ctx = (x, st1, h2, st2)
        return out

    def backward(self, dout):
        x, st1, h2, st2 = self.ctx
        self.ctx = None
        B, H, W, Cin = x.shape
        dh3 = self.c2.dgrad(dout)
        dh2 = ops.groupnorm_bwd(dh3, h2, *self.n2, st2, self.G, self.eps, True, arena=self.arena)
        dh1 = self.c1.dgrad(dh2)
        if self.sc is None:
            return ops.groupnorm_bwd(dh1, x, *self.n1, st1, self.G, self.eps, True, add=dout, arena=self.arena)
        dx = ops.groupnorm_bwd(dh1, x, *self.n1, st1, self.G, self.eps, True, arena=self.arena)
        return self.sc.dgrad(dout.view(B * H * W, -1), residual=dx.view(-1, Cin)).view(B, H, W, Cin)


class _Transformer:
    """Transformer2DModel with one BasicTransformerBlock."""
    arena = None  # see _Resnet.arena

    def __init__(self, sd, p, cfg: UNetConfig, heads: int):
        self.G, self.heads = cfg.norm_num_groups, heads
        self.prefix = p  # diffusers module path of the Transformer2DModel ("down_blocks.0.attentions.0.")
        self.norm = (sd[p + "norm.weight"], sd[p + "norm.bias"])
        self.pi = _Linear(sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
        self.po = _Linear(sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
        t = p + "transformer_blocks.0."
        self.ln1 = (sd[t + "norm1.weight"], sd[t + "norm1.bias"])
        self.ln2 = (sd[t + "norm2.weight"], sd[t + "norm2.bias"])
        self.ln3 = (sd[t + "norm3.weight"], sd[t + "norm3.bias"])
        self.qkv = _Linear(torch.cat([sd[t + "attn1.to_q.weight"], sd[t + "attn1.to_k.weight"],
                                      sd[t + "attn1.to_v.weight"]], 0))
        self.o1 = _Linear(sd[t + "attn1.to_out.0.weight"], sd[t + "attn1.to_out.0.bias"])
        self.q2 = _Linear(sd[t + "attn2.to_q.weight"])
        self.kv2 = _Linear(torch.cat([sd[t + "attn2.to_k.weight"], sd[t + "attn2.to_v.weight"]], 0))
        self.o2 = _Linear(sd[t + "attn2.to_out.0.weight"], sd[t + "attn2.to_out.0.bias"])
        self.ff1 = _Linear(sd[t + "ff.net.0.proj.weight"], sd[t + "ff.net.0.proj.bias"])
        self.ff2 = _Linear(sd[t + "ff.net.2.weight"], sd[t + "ff.net.2.bias"])
        self.first = False  # first cross-attention of the net: backward stops at dK/dV
        self.ehs_ready = None  # set on the first cross-attention: event after which encoder_hidden_states is valid

    def forward(self, x, ehs, save, kv2=None):
        """kv2: this block's [to_k | to_v](encoder_hidden_states) [B, L, 2C], normally a column slice of the ONE GEMM
        that projects the text states for all cross-attentions of the net (UNetEngine.forward)."""
        B, H, W, Cc = x.shape
        N, M = H * W, B * H * W
        L = ehs.shape[1]
        hn, st0 = ops.groupnorm(x, *self.norm, self.G, 1e-6, False, arena=self.arena)
        h0 = self.pi.fwd(hn.view(M, Cc))
        n1, s1 = ops.layernorm(h0, *self.ln1)
        qkv = self.qkv.fwd(n1).view(B, N, 3 * Cc)
        o1, lse1 = ops.attn_fwd(qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:], self.heads)
        h1 = self.o1.fwd(o1.view(M, Cc), residual=h0)
        n2, s2 = ops.layernorm(h1, *self.ln2)
        q2 = self.q2.fwd(n2).view(B, N, Cc)
        if callable(kv2):
            kv2 = kv2(self)
        elif kv2 is None:
            if self.ehs_ready is not None:  # the text encoder runs on another stream up to here (trainer.py)
                torch.cuda.current_stream().wait_event(self.ehs_ready)
            kv2 = self.kv2.fwd(ehs.reshape(B * L, -1)).view(B, L, 2 * Cc)
        o2, lse2 = ops.attn_fwd(q2, kv2[..., :Cc], kv2[..., Cc:], self.heads)
        h2 = self.o2.fwd(o2.view(M, Cc), residual=h1)
        n3, s3 = ops.layernorm(h2, *self.ln3)
        f1 = self.ff1.fwd(n3)
        g = ops.geglu(f1)
        h3 = self.ff2.fwd(g, residual=h2)
        out = self.po.fwd(h3, residual=x.view(M, Cc)).view(B, H, W, Cc)
        if save:
            self.ctx = (x, st0, h0, s1, qkv, o1, lse1, h1, s2, q2, kv2, o2, lse2, h2, s3, f1)
        return out

    def backward(self, dout, d_ehs, dkv2=None):
        """dkv2 (fp16 [B, L, 2C] view): where d[to_k | to_v] goes when the caller projects all blocks' K/V gradients
        back to the text states with one GEMM at the end; None: projected and accumulated into d_ehs here."""
        (x, st0, h0, s1, qkv, o1, lse1, h1, s2, q2, kv2, o2, lse2, h2, s3, f1) = self.ctx
        self.ctx = None
        B, H, W, Cc = x.shape
        N, M = H * W, B * H * W
        L = kv2.shape[1]
        dout2 = dout.view(M, Cc)
        dh3 = self.po.dgrad(dout2)
        dg = self.ff2.dgrad(dh3)
        df1 = ops.geglu_bwd(dg, f1)
        dn3 = self.ff1.dgrad(df1)
        dh2 = ops.layernorm_bwd(dn3, h2, self.ln3[0], s3, add=dh3)
        # cross attention
        do2 = self.o2.dgrad(dh2).view(B, N, Cc)
        batched = dkv2 is not None
        if not batched:
            dkv2 = torch.empty((B, L, 2 * Cc), device=kv2.device, dtype=POLICY.act)
        # 77 text tokens = a single KV tile: dQ leaves the kernel as fp16 (no fp32 accumulator, memset or cast)
        one_tile = L <= 128
        dq2, _, _ = ops.attn_bwd(q2, kv2[..., :Cc], kv2[..., Cc:], o2, do2, lse2, self.heads,
                                 need_dq=not self.first, dk=dkv2[..., :Cc], dv=dkv2[..., Cc:],
                                 dq_out=(one_tile and not self.first) or None)
        if not batched:
            ops.gemm(dkv2.view(B * L, 2 * Cc), self.kv2.wt, out=d_ehs, out_kind=C.TB_OUT_F32_ACC)
        if self.first:
            return None
        dq2h = dq2.view(M, Cc) if one_tile else ops.cast_f32_f16(dq2.view(M, Cc))
        dn2 = self.q2.dgrad(dq2h)
        dh1 = ops.layernorm_bwd(dn2, h1, self.ln2[0], s2, add=dh2)
        # self attention
        do1 = self.o1.dgrad(dh1).view(B, N, Cc)
        dqkv = torch.empty_like(qkv)
        dq1, _, _ = ops.attn_bwd(qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:], o1, do1, lse1,
                                 self.heads, dk=dqkv[..., Cc:2 * Cc], dv=dqkv[..., 2 * Cc:])
        ops.cast_f32_f16(dq1.view(M, Cc), out=dqkv.view(M, 3 * Cc)[:, :Cc])
        dn1 = self.qkv.dgrad(dqkv.view(M, 3 * Cc))
        dh0 = ops.layernorm_bwd(dn1, h0, self.ln1[0], s1, add=dh1)
        dhn = self.pi.dgrad(dh0).view(B, H, W, Cc)
        return ops.groupnorm_bwd(dhn, x, *self.norm, st0, self.G, 1e-6, False, add=dout, arena=self.arena)


class _Down:
    def __init__(self, sd, p):
        self.w = _Conv3(sd[p + "conv.weight"], sd[p + "conv.bias"])

    def forward(self, x):
        B, H, W, Cc = x.shape
        col = ops.im2col3x3s2(x)
        return ops.gemm(col, self.w.wk, bias=self.w.b).view(B, H // 2, W // 2, -1)

    def backward(self, dout):
        return self.w.dgrad(ops.zero_stuff2x(dout))


class _Up:
    def __init__(self, sd, p):
        self.w = _Conv3(sd[p + "conv.weight"], sd[p + "conv.bias"])

    def forward(self, x):
        return self.w.fwd(ops.upsample2x(x))

    def backward(self, dout):
        return ops.upsample2x_bwd(self.w.dgrad(dout))


class CrossKVLora:
    """The UNet adapter of ``--unet_params_to_train crossattn_kv`` (train_textboost.py:712-721): peft LoRA
    (``LoraConfig(r, lora_alpha=r, init_lora_weights="gaussian", target_modules=["attn2.to_k", "attn2.to_v"])``) on the
    K and V projections of every cross-attention.  Adapter 2i / 2i+1 is to_k / to_v of the i-th transformer block in
    forward order; they are stacked the way the engine stacks the projections themselves (one GEMM for all blocks):

      params = [ A (n_adapters*r, ctx) | B (KV, r) ]  fp32 master, flat (the third optimiser group, :838-841);
      grads  = same layout, accumulated by backward (loss-scaled like every other gradient of the step).

    peft's gaussian init: lora_A ~ N(0, 1/r), lora_B = 0; scaling = lora_alpha / r = 1."""

    def __init__(self, engine: "UNetEngine", r: int, alpha: Optional[float] = None, seed: int = 0):
        if not 1 <= r <= 16:
            raise ValueError(f"UNet LoRA rank {r}: 1..16")
        dev = engine._kv_w.device
        self.r, self.scaling = r, (alpha if alpha is not None else r) / r
        self.ctx, self.KV = engine._kv_w.shape[1], engine._kv_total
        self.n_adapters = 2 * len(engine._attns)
        self.R = self.n_adapters * r
        self.n_a, self.n_b = self.R * self.ctx, self.KV * r
        self.params = torch.zeros(self.n_a + self.n_b, device=dev, dtype=F32)
        self.grads = torch.zeros_like(self.params)
        g = torch.Generator().manual_seed(seed)
        self.A().copy_((torch.randn(self.R, self.ctx, generator=g) / r).to(dev))
        blk = torch.empty(self.KV, dtype=torch.int32)
        off = [0]
        self.names = []  # diffusers / peft module path of every adapter
        for i, a in enumerate(engine._attns):
            Cc = a.kv2.w.shape[0] // 2
            assert Cc % 8 == 0 and a.kv_off % 8 == 0  # the kernels walk K/V in 8-column vectors, one adapter each
            blk[a.kv_off:a.kv_off + Cc] = 2 * i
            blk[a.kv_off + Cc:a.kv_off + 2 * Cc] = 2 * i + 1
            off += [a.kv_off + Cc, a.kv_off + 2 * Cc]
            self.names += [a.prefix + "transformer_blocks.0.attn2.to_k", a.prefix + "transformer_blocks.0.attn2.to_v"]
        self.blk, self.off = blk.to(dev), torch.tensor(off, dtype=torch.int32, device=dev)
        self._ctx = None

    def A(self, buf=None):
        return (self.params if buf is None else buf)[:self.n_a].view(self.R, self.ctx)

    def B(self, buf=None):
        return (self.params if buf is None else buf)[self.n_a:].view(self.KV, self.r)

    def forward(self, ehs2d, kv2d, save):
        Z = ops.unet_lora_fwd(ehs2d, self.A(), self.B(), self.blk, kv2d, self.r, self.scaling)
        self._ctx = (ehs2d, Z) if save else None

    def backward(self, dkv2d, d_ehs2d):
        ehs2d, Z = self._ctx
        self._ctx = None
        ops.unet_lora_bwd(dkv2d, ehs2d, self.A(), self.B(), Z, self.blk, self.off, self.A(self.grads),
                          self.B(self.grads), d_ehs2d, self.r, self.scaling)

    # peft state-dict surface (get_peft_model_state_dict names: "<module>.lora_A.weight" [r, ctx], ".lora_B.weight" [C, r])
    def state_dict(self):
        sd = {}
        for i, n in enumerate(self.names):
            lo, hi = int(self.off[i]), int(self.off[i + 1])
            sd[n + ".lora_A.weight"] = self.A()[i * self.r:(i + 1) * self.r].detach().clone()
            sd[n + ".lora_B.weight"] = self.B()[lo:hi].detach().clone()
        return sd

    def load_state_dict(self, sd):
        off = self.off.tolist()
        for i, n in enumerate(self.names):
            self.A()[i * self.r:(i + 1) * self.r].copy_(sd[n + ".lora_A.weight"])
            self.B()[off[i]:off[i + 1]].copy_(sd[n + ".lora_B.weight"])


class UNetEngine:
    """Weights + forward/backward orchestration.  `sd` maps diffusers keys to fp16 CUDA tensors."""

    def __init__(self, cfg: UNetConfig, sd: Dict[str, torch.Tensor]):
        self.cfg = cfg
        ch = cfg.block_out_channels
        sd = {k: v.detach().to(dtype=POLICY.act).contiguous() for k, v in sd.items()}
        self.conv_in_w, self.conv_in_b = sd["conv_in.weight"], sd["conv_in.bias"]
        self.t1 = _Linear(sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])
        self.t2 = _Linear(sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])
        self.down: List[dict] = []
        for i in range(len(ch)):
            p = f"down_blocks.{i}."
            blk = {"res": [], "attn": [], "down": None}
            for j in range(cfg.layers_per_block):
                blk["res"].append(_Resnet(sd, f"{p}resnets.{j}.", cfg))
                if cfg.down_has_attn[i]:
                    blk["attn"].append(_Transformer(sd, f"{p}attentions.{j}.", cfg, cfg.attention_head_dim[i]))
            if i != len(ch) - 1:
                blk["down"] = _Down(sd, f"{p}downsamplers.0.")
            self.down.append(blk)
        self.mid_res = [_Resnet(sd, "mid_block.resnets.0.", cfg), _Resnet(sd, "mid_block.resnets.1.", cfg)]
        self.mid_attn = _Transformer(sd, "mid_block.attentions.0.", cfg, cfg.attention_head_dim[-1])
        self.up: List[dict] = []
        rev_heads = list(reversed(cfg.attention_head_dim))
        rev_attn = list(reversed(cfg.down_has_attn))
        for i in range(len(ch)):
            p = f"up_blocks.{i}."
            blk = {"res": [], "attn": [], "up": None}
            for j in range(cfg.layers_per_block + 1):
                blk["res"].append(_Resnet(sd, f"{p}resnets.{j}.", cfg))
                if rev_attn[i]:
                    blk["attn"].append(_Transformer(sd, f"{p}attentions.{j}.", cfg, rev_heads[i]))
            if i != len(ch) - 1:
                blk["up"] = _Up(sd, f"{p}upsamplers.0.")
            self.up.append(blk)
        self.norm_out = (sd["conv_norm_out.weight"], sd["conv_norm_out.bias"])
        self.conv_out_w, self.conv_out_b = sd["conv_out.weight"], sd["conv_out.bias"]
        # the first cross-attention in forward order: nothing upstream of it depends on the text
        first = None
        for blk in self.down:
            if blk["attn"]:
                first = blk["attn"][0]
                break
        self.first_attn = first if first is not None else self.mid_attn
        self.first_attn.first = True
        self._saved = None
        # Horizontal batching of the launches that depend only on the step's inputs (every GEMM with 8 or 616 rows is
        # pure latency: ~8-10 us per launch for a few MFLOP):
        #   * time_emb_proj of all ResnetBlock2D: ONE [B,1280] x [sum Cout,1280]^T GEMM; blocks read column slices
        #   * attn2.to_k / to_v of all Transformer2D blocks: ONE [B*77, ctx] x [sum 2C, ctx]^T GEMM, and in the
        #     backward ONE [B*77, sum 2C] x [ctx, sum 2C]^T GEMM for d(encoder_hidden_states)
        self._resnets = [r for blk in self.down for r in blk["res"]] + self.mid_res + \
                        [r for blk in self.up for r in blk["res"]]
        self._temb_w = torch.cat([r.t.w for r in self._resnets], 0).contiguous()
        self._temb_b = torch.cat([r.t.b for r in self._resnets], 0).contiguous()
        off = 0
        for r in self._resnets:
            r.t_off, off = off, off + r.t.w.shape[0]
        self._attns = [a for blk in self.down for a in blk["attn"]] + [self.mid_attn] + \
                      [a for blk in self.up for a in blk["attn"]]
        self._kv_w = torch.cat([a.kv2.w for a in self._attns], 0).contiguous()       # [sum 2C, ctx]
        self._kv_wt = self._kv_w.t().contiguous()                                    # [ctx, sum 2C]
        off = 0
        for a in self._attns:
            a.kv_off, off = off, off + a.kv2.w.shape[0]
        self._kv_total = off
        for a in self._attns:  # the per-block copies are only kept for the unbatched entry points (tests)
            a.kv2.wt = self._kv_wt[:, a.kv_off:a.kv_off + a.kv2.w.shape[0]]
        self.kv_lora: Optional[CrossKVLora] = None

    def add_cross_kv_lora(self, r: int, alpha: Optional[float] = None, seed: int = 0) -> CrossKVLora:
        """unet.add_adapter(LoraConfig(target_modules=["attn2.to_k", "attn2.to_v"])) of train_textboost.py:712-720."""
        self.kv_lora = CrossKVLora(self, r, alpha, seed)
        return self.kv_lora

    def _set_arena(self, arena):
        self._arena = arena
        for m in self._resnets + self._attns:
            m.arena = arena

    # ------------------------------------------------------------------ forward
    def forward(self, sample, timesteps, ehs, save_for_backward=True, ehs_ready=None):
        """sample [B,4,H,W] fp16 NCHW, timesteps int64 [B], ehs [B,L,ctx] fp16 -> [B,4,H,W] fp16.

        ehs_ready (torch.cuda.Event): `ehs` is still being produced on another stream; nothing up to the first
        cross-attention K/V projection reads it, so only that point waits for the event -- conv_in, the time MLP, the
        first resnet and the first self-attention overlap the text encoder (the reference runs them back to back,
        train_textboost.py:1054-1067)."""
        cfg = self.cfg
        self.first_attn.ehs_ready = ehs_ready
        assert sample.dtype == POLICY.act and ehs.dtype == POLICY.act and timesteps.dtype == torch.int64
        sample = sample.contiguous()
        ehs = ehs.contiguous()
        save = save_for_backward
        te = ops.timestep_embedding(timesteps, cfg.block_out_channels[0])
        e1 = self.t1.fwd(te, act=C.TB_ACT_SILU)
        temb_act = self.t2.fwd(e1, act=C.TB_ACT_SILU)  # every consumer applies SiLU first
        tp_all = ops.gemm(temb_act, self._temb_w, bias=self._temb_b)  # every block's time_emb_proj at once
        n_gn = 2 * len(self._resnets) + len(self._attns) + 1
        self._set_arena(ops.StatsArena(n_gn, sample.shape[0], cfg.norm_num_groups, sample.device))

        def tp_of(r):
            return tp_all[:, r.t_off:r.t_off + r.t.w.shape[0]]

        B, L = ehs.shape[0], ehs.shape[1]
        kv_all = [None]

        def kv_of(a):
            if kv_all[0] is None:  # first cross-attention: wait for the text encoder, project K/V for all 16 blocks
                if ehs_ready is not None:
                    torch.cuda.current_stream().wait_event(ehs_ready)
                ehs2d = ehs.reshape(B * L, -1)
                kv2d = ops.gemm(ehs2d, self._kv_w)
                if self.kv_lora is not None:
                    self.kv_lora.forward(ehs2d, kv2d, save_for_backward)
                kv_all[0] = kv2d.view(B, L, self._kv_total)
            return kv_all[0][..., a.kv_off:a.kv_off + a.kv2.w.shape[0]]

        x = ops.conv_in(sample, self.conv_in_w, self.conv_in_b)
        skips = [x]
        for blk in self.down:
            for j, r in enumerate(blk["res"]):
                x = r.forward(x, temb_act, save, tp_of(r))
                if blk["attn"]:
                    x = blk["attn"][j].forward(x, ehs, save, kv_of)
                skips.append(x)
            if blk["down"] is not None:
                x = blk["down"].forward(x)
                skips.append(x)
        x = self.mid_res[0].forward(x, temb_act, save, tp_of(self.mid_res[0]))
        x = self.mid_attn.forward(x, ehs, save, kv_of)
        x = self.mid_res[1].forward(x, temb_act, save, tp_of(self.mid_res[1]))
        cat_split = []
        for blk in self.up:
            for j, r in enumerate(blk["res"]):
                s = skips.pop()
                cat_split.append(x.shape[-1])
                x = ops.concat_channels(x, s)
                x = r.forward(x, temb_act, save, tp_of(r))
                if blk["attn"]:
                    x = blk["attn"][j].forward(x, ehs, save, kv_of)
            if blk["up"] is not None:
                x = blk["up"].forward(x)
        hn, st = ops.groupnorm(x, *self.norm_out, cfg.norm_num_groups, cfg.norm_eps, True, arena=self._arena)
        out = ops.conv_out(hn, self.conv_out_w, self.conv_out_b)
        self._set_arena(None)
        self.first_attn.ehs_ready = None
        if save:
            self._saved = (x, st, cat_split, ehs.shape)
        return out

    # ------------------------------------------------------------------ backward (dgrad to ehs only)
    def backward(self, dout, d_ehs: Optional[torch.Tensor] = None):
        """dout [B,4,H,W] fp16 -> d(encoder_hidden_states) fp32 [B,L,ctx] (accumulated if given)."""
        cfg = self.cfg
        assert self._saved is not None, "forward(save_for_backward=True) must precede backward"
        x_last, st, cat_split, ehs_shape = self._saved
        self._saved = None
        B, L, ctx = ehs_shape
        if d_ehs is None:
            d_ehs = ops.zeros((B * L, ctx), dout.device)
        d2 = d_ehs.view(B * L, ctx)
        n_gn = 2 * len(self._resnets) + len(self._attns) + 1
        self._set_arena(ops.StatsArena(n_gn, dout.shape[0], cfg.norm_num_groups, dout.device))
        dh = ops.conv_out_bwd(dout.contiguous(), self.conv_out_w)
        dx = ops.groupnorm_bwd(dh, x_last, *self.norm_out, st, cfg.norm_num_groups, cfg.norm_eps, True,
                               arena=self._arena)
        dskips = []  # gradients of the skip tensors in the order they are popped in forward
        # d[to_k | to_v] of all cross-attentions side by side: one GEMM at the end takes them back to the text states
        dkv_all = torch.empty((B, L, self._kv_total), device=dout.device, dtype=POLICY.act)

        def dkv_of(a):
            return dkv_all[..., a.kv_off:a.kv_off + a.kv2.w.shape[0]]

        ci = len(cat_split)
        for blk in reversed(self.up):
            if blk["up"] is not None:
                dx = blk["up"].backward(dx)
            for j in reversed(range(len(blk["res"]))):
                if blk["attn"]:
                    dx = blk["attn"][j].backward(dx, d2, dkv_of(blk["attn"][j]))
                dcat = blk["res"][j].backward(dx)
                ci -= 1
                dx, ds = ops.split_channels(dcat, cat_split[ci])
                dskips.append(ds)
        # forward pushes skips s0..s(n-1) and the up path pops them last-first, so walking the up path
        # backwards yields d(s0), d(s1), ...: dskips[k] is the gradient of the k-th pushed skip.
        n = len(dskips)
        dskip_of = dskips
        dx = self.mid_res[1].backward(dx)
        dx = self.mid_attn.backward(dx, d2, dkv_of(self.mid_attn))
        dx = self.mid_res[0].backward(dx)
        k = n - 1  # index of the skip produced last in forward order
        done = False
        for blk in reversed(self.down):
            if blk["down"] is not None:
                ops.copy2d(dx.view(-1, dx.shape[-1]), dskip_of[k].reshape(-1, dx.shape[-1]), accumulate=True)
                k -= 1
                dx = blk["down"].backward(dx)
            for j in reversed(range(len(blk["res"]))):
                ops.copy2d(dx.view(-1, dx.shape[-1]), dskip_of[k].reshape(-1, dx.shape[-1]), accumulate=True)
                k -= 1
                if blk["attn"]:
                    a = blk["attn"][j]
                    dx = a.backward(dx, d2, dkv_of(a))
                    if a.first:
                        done = True
                        break
                dx = blk["res"][j].backward(dx)
            if done:
                break
        ops.gemm(dkv_all.view(B * L, self._kv_total), self._kv_wt, out=d2, out_kind=C.TB_OUT_F32_ACC)
        if self.kv_lora is not None:
            self.kv_lora.backward(dkv_all.view(B * L, self._kv_total), d2)
        self._set_arena(None)
        # drop contexts of the pruned prefix
        for blk in self.down:
            for r in blk["res"]:
                r.ctx = None
        return d_ehs.view(B, L, ctx)
