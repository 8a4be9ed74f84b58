"""CLIP text encoder (ViT-L/14 text tower, or OpenCLIP-H) with rank-r LoRA fused into the QKV GEMM,
forward and hand-derived backward on the B200 kernels.

Replaces, for the TextBoost path, ``transformers.CLIPTextModel`` + peft LoRA ``Linear`` +
``textboost.text_encoder.TextBoostModel`` (/root/reference/textboost/text_encoder.py:17-87, called from
train_textboost.py:1054-1059 and :1099-1100; LoRA configured at :702-709).

Precision policy mirrors accelerate fp16 autocast: master weights fp32, residual
stream / LayerNorm / softmax statistics fp32, GEMM operands fp16 (static fp16 copies of the frozen base
weights), fp32 accumulation.

LoRA is not a separate pair of tiny GEMMs: with xa = LN(x) A^T (r columns per target) appended to the
GEMM's K dimension and scaling*B appended to the weight, one tcgen05 GEMM computes
  [q|k|v] = [LN(x) | xa] [W | sB]^T + b
and the same trick with the transposed operand gives dLN(x) and dxa in one dgrad GEMM.  The reference's
targets are q/k/v (train_textboost.py:705); out_proj (BASELINE.json's "QKV/out projections", the first entry
of the reference's commented-out wider list, :702-708) extends the out-projection GEMM the same way:
  x2 = [o | o A_o^T] [W_o | s B_o]^T + b_o + x.    Any subset of {q,k,v,out}_proj, rank 1..16.

Trainable state lives in ONE flat fp32 buffer  [A (layers x T r x D) | B (layers x T x D x r) | added rows]
(T targets in the order q, k, v, out) with a same-shaped gradient buffer: that buffer is what the single NCCL
all-reduce and the fused AdamW see.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, Optional

import torch

from . import _cabi as C
from . import ops

from .precision import POLICY
F32 = torch.float32
EOS_ID = 49407  # hard-coded in textboost/text_encoder.py:71
RMAX = 16       # largest LoRA rank (K extension = targets x r columns, rounded up to 16, <= 64)


@dataclasses.dataclass
class ClipConfig:
    vocab_size: int = 49408
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    max_position_embeddings: int = 77
    hidden_act: str = "quick_gelu"
    layer_norm_eps: float = 1e-5

    @staticmethod
    def clip_l():
        return ClipConfig()

    @staticmethod
    def openclip_h():
        return ClipConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                          num_attention_heads=16, hidden_act="gelu")


LORA_TARGETS = ("q_proj", "k_proj", "v_proj")                  # the reference's configuration (:705)
SUPPORTED_TARGETS = ("q_proj", "k_proj", "v_proj", "out_proj")  # canonical order of the trainable state


def canonical_targets(targets) -> tuple:
    targets = tuple(targets)
    bad = [t for t in targets if t not in SUPPORTED_TARGETS]
    if bad or not targets or len(set(targets)) != len(targets):
        raise NotImplementedError(
            f"LoRA target_modules={list(targets)}: the fused path covers any subset of {list(SUPPORTED_TARGETS)} "
            "(fc1 / fc2 of the reference's commented-out list, train_textboost.py:702-708, are not built)")
    return tuple(t for t in SUPPORTED_TARGETS if t in targets)


def _pad16(n: int) -> int:
    return (n + 15) // 16 * 16


class TrainableState:
    """Flat fp32 parameter / gradient buffers shared by the encoder, the all-reduce and the optimiser."""

    def __init__(self, n_layers: int, D: int, r: int, n_rows: int, device, n_targets: int = 3):
        self.n_layers, self.D, self.r, self.n_rows = n_layers, D, r, n_rows
        self.T = n_targets if r > 0 else 0
        self.n_a = n_layers * self.T * r * D
        self.n_b = n_layers * self.T * D * r
        self.n_lora = self.n_a + self.n_b
        self.n_total = self.n_lora + n_rows * D
        self.params = torch.zeros(self.n_total, device=device, dtype=F32)
        self.grads = torch.zeros(self.n_total, device=device, dtype=F32)

    def A(self, l, buf=None):  # [T*r, D]
        buf = self.params if buf is None else buf
        sz = self.T * self.r * self.D
        return buf[l * sz:(l + 1) * sz].view(self.T * self.r, self.D)

    def B(self, l, buf=None):  # [T, D, r]
        buf = self.params if buf is None else buf
        sz = self.T * self.D * self.r
        return buf[self.n_a + l * sz:self.n_a + (l + 1) * sz].view(self.T, self.D, self.r)

    def rows(self, buf=None):  # [n_rows, D]
        buf = self.params if buf is None else buf
        return buf[self.n_lora:].view(self.n_rows, self.D)

    def b_segment(self, buf=None):
        buf = self.params if buf is None else buf
        return buf[self.n_a:self.n_lora]


class ClipEngine:
    """sd: HF ``text_model.*`` keys (fp32).  LoRA keys in peft naming are honoured if present."""

    def __init__(self, cfg: ClipConfig, sd: Dict[str, torch.Tensor], device, lora_r: int = 0,
                 lora_alpha: Optional[int] = None, n_base: Optional[int] = None, seed: int = 0,
                 lora_targets=LORA_TARGETS):
        self.cfg = cfg
        self.device = torch.device(device)
        D, nl = cfg.hidden_size, cfg.num_hidden_layers
        self.D, self.nl, self.heads = D, nl, cfg.num_attention_heads
        assert D // self.heads == 64, "CLIP text towers use head_dim 64"
        self.r = lora_r
        self.scaling = (lora_alpha if lora_alpha is not None else lora_r) / lora_r if lora_r else 0.0
        self.act = C.TB_ACT_QUICK_GELU if cfg.hidden_act == "quick_gelu" else C.TB_ACT_GELU
        if not 0 <= lora_r <= RMAX:
            raise NotImplementedError(f"LoRA rank {lora_r}: the K-extension path covers ranks 1..{RMAX}")
        self.targets = canonical_targets(lora_targets) if lora_r else ()
        qkv_t = [t for t in self.targets if t != "out_proj"]
        self.Tq = len(qkv_t)                                  # LoRA targets inside the fused QKV GEMM
        self.qkv_mask = sum(1 << LORA_TARGETS.index(t) for t in qkv_t)
        self.has_o = "out_proj" in self.targets
        self.Rq, self.Ro = _pad16(self.Tq * lora_r), (_pad16(lora_r) if self.has_o else 0)
        self.Kext = D + self.Rq                               # K of the fused QKV GEMM
        self.Ko = D + self.Ro                                 # K of the out-projection GEMM

        def g32(k):
            return sd[k].detach().to(device=self.device, dtype=F32).contiguous()

        def g16(k):
            return sd[k].detach().to(device=self.device, dtype=POLICY.act).contiguous()

        def base_key(prefix, kind):  # peft renames W to base_layer.W after injection
            k = f"{prefix}.base_layer.{kind}"
            return k if k in sd else f"{prefix}.{kind}"

        emb = g32("text_model.embeddings.token_embedding.weight")
        self.n_base = emb.shape[0] if n_base is None else n_base
        self.tok_base = emb[:self.n_base].contiguous()
        n_rows = emb.shape[0] - self.n_base
        self.pos = g32("text_model.embeddings.position_embedding.weight")
        self.state = TrainableState(nl, D, lora_r, n_rows, self.device, len(self.targets))
        if n_rows:
            self.state.rows().copy_(emb[self.n_base:])
        self.layers = []
        gen = torch.Generator().manual_seed(seed)
        for l in range(nl):
            p = f"text_model.encoder.layers.{l}."
            L = {}
            L["ln1"] = (g32(p + "layer_norm1.weight"), g32(p + "layer_norm1.bias"))
            L["ln2"] = (g32(p + "layer_norm2.weight"), g32(p + "layer_norm2.bias"))
            ws, bs = [], []
            for t in LORA_TARGETS:
                ws.append(g16(base_key(p + "self_attn." + t, "weight")))
                bs.append(g16(base_key(p + "self_attn." + t, "bias")))
            wqkv = torch.cat(ws, 0)  # [3D, D]
            wext = torch.zeros((3 * D, self.Kext), device=self.device, dtype=POLICY.act)
            wext[:, :D] = wqkv
            wext_t = torch.zeros((self.Kext, 3 * D), device=self.device, dtype=POLICY.act)
            wext_t[:D] = wqkv.t()
            L["wqkv"], L["wqkv_t"], L["bqkv"] = wext, wext_t, torch.cat(bs, 0)
            for name, key in (("o", "self_attn.out_proj"), ("f1", "mlp.fc1"), ("f2", "mlp.fc2")):
                w = g16(base_key(p + key, "weight"))
                L["w" + name], L["w" + name + "_t"] = w, w.t().contiguous()
                L["b" + name] = g16(base_key(p + key, "bias"))
            if self.has_o:  # [W_o | s B_o] and its transpose, extension filled by pack_lora
                wo = torch.zeros((D, self.Ko), device=self.device, dtype=POLICY.act)
                wo[:, :D] = L["wo"]
                wo_t = torch.zeros((self.Ko, D), device=self.device, dtype=POLICY.act)
                wo_t[:D] = L["wo_t"]
                L["wo"], L["wo_t"] = wo, wo_t
            self.layers.append(L)
            if lora_r:
                for ti, t in enumerate(self.targets):
                    ka = f"{p}self_attn.{t}.lora_A.default.weight"
                    kb = f"{p}self_attn.{t}.lora_B.default.weight"
                    if ka in sd:
                        self.state.A(l)[ti * lora_r:(ti + 1) * lora_r].copy_(sd[ka].to(self.device, F32))
                        self.state.B(l)[ti].copy_(sd[kb].to(self.device, F32))
                    else:  # peft init_lora_weights="gaussian": A ~ N(0, 1/r), B = 0
                        a = torch.randn((lora_r, D), generator=gen) / lora_r
                        self.state.A(l)[ti * lora_r:(ti + 1) * lora_r].copy_(a.to(self.device))
        self.lnf = (g32("text_model.final_layer_norm.weight"), g32("text_model.final_layer_norm.bias"))
        self.null_embedding = torch.zeros((cfg.max_position_embeddings, D), device=self.device, dtype=F32)
        self.use_fixed_special = False
        self.null_override = True  # False: plain CLIPTextModel forward (inference.py's pipeline), inference only
        self.decay = torch.ones(1, device=self.device, dtype=F32)  # lazy weight decay of frozen rows (D8)
        self._ctx = []  # saved forward contexts, most recent last (one per forward awaiting its backward)

    # textboost/text_encoder.py:28-32
    def set_null_embedding(self, t: torch.Tensor):
        assert t.shape == self.null_embedding.shape, (t.shape, self.null_embedding.shape)
        self.null_embedding = t.detach().to(self.device, F32).contiguous()
        self.use_fixed_special = True

    def pack_lora(self):
        """refresh the B-dependent extension blocks of the fused QKV weights (after every optimiser step)."""
        if not self.r:
            return
        st = self.state
        for l, L in enumerate(self.layers):
            if self.Tq:
                C.call("tb_lora_pack", C.ptr(st.B(l)), C.ptr(L["wqkv"]), C.ptr(L["wqkv_t"]), 3, self.qkv_mask,
                       self.D, self.r, self.Rq, self.scaling, C.stream_ptr())
            if self.has_o:
                C.call("tb_lora_pack", C.ptr(st.B(l)[self.Tq]), C.ptr(L["wo"]), C.ptr(L["wo_t"]), 1, 1, self.D,
                       self.r, self.Ro, self.scaling, C.stream_ptr())

    # ------------------------------------------------------------------ forward
    def forward(self, input_ids: torch.Tensor, save_for_backward: bool = False) -> torch.Tensor:
        """input_ids int64 [B, L] on the device -> last_hidden_state fp32 [B, L, D] (after the override)."""
        assert input_ids.dtype == torch.int64 and input_ids.device.type == self.device.type
        ids = input_ids.contiguous()
        B, Lq = ids.shape
        D, M = self.D, B * Lq
        st = self.state
        s = C.stream_ptr()
        x = torch.empty((M, D), device=self.device, dtype=F32)
        C.call("tb_clip_embed", C.ptr(ids), C.ptr(self.tok_base), C.ptr(st.rows() if st.n_rows else None),
               C.ptr(self.decay), C.ptr(self.pos), C.ptr(x), M, Lq, D, self.n_base, st.n_rows, s)
        saved = []
        for l, L in enumerate(self.layers):
            y_ext = torch.empty((M, self.Kext), device=self.device, dtype=POLICY.act)
            if self.Tq:  # LayerNorm and the LoRA down-projection [LN(x) | LN(x) A^T] in one launch
                st1 = ops.layernorm_lora_fwd(x, *L["ln1"], st.A(l)[:self.Tq * self.r], y_ext, self.Rq,
                                             eps=self.cfg.layer_norm_eps)
            else:
                _, st1 = ops.layernorm(x, *L["ln1"], eps=self.cfg.layer_norm_eps, out=y_ext[:, :D])
            qkv = ops.gemm(y_ext, L["wqkv"], bias=L["bqkv"])
            # causal softmax(QK^T/sqrt(64))V on the tcgen05 flash kernels, reading q/k/v in place from the fused
            # projection output (transformers CLIPAttention under the causal mask, text_encoder.py:62-69)
            q3 = qkv.view(B, Lq, 3 * D)
            # the attention writes straight into the first D columns of the out-projection's (extended) A operand
            o = torch.empty((M, self.Ko), device=self.device, dtype=POLICY.act)
            _, lse = ops.attn_fwd(q3[..., :D], q3[..., D:2 * D], q3[..., 2 * D:], self.heads, causal=True,
                                  out=o.view(B, Lq, self.Ko)[..., :D])
            if self.has_o:
                C.call("tb_lora_down", C.ptr(o), self.Ko, C.ptr(st.A(l)[self.Tq * self.r:]), M, D, self.r, self.Ro, s)
            x2 = ops.gemm(o, L["wo"], bias=L["bo"], residual=x, out_kind=C.TB_OUT_F32)
            y2, st2 = ops.layernorm(x2, *L["ln2"], eps=self.cfg.layer_norm_eps)
            # (activation NOT fused into the fc1 epilogue: measured, the per-element exp / erf in the GEMM's 8 epilogue
            # warps costs ~11 us per launch against ~4 us for the separate full-occupancy kernel)
            u = ops.gemm(y2, L["wf1"], bias=L["bf1"])
            a = torch.empty_like(u)
            C.call("tb_act_fwd_f16", C.ptr(u), C.ptr(a), u.numel(), self.act, s)
            x3 = ops.gemm(a, L["wf2"], bias=L["bf2"], residual=x2, out_kind=C.TB_OUT_F32)
            if save_for_backward:
                saved.append((x, st1, y_ext, qkv, x2, st2, u, o, lse))
            x = x3
        out, stf = ops.layernorm(x, *self.lnf, eps=self.cfg.layer_norm_eps, out_dtype=F32)
        if self.null_override:
            C.call("tb_null_override", C.ptr(ids), C.ptr(self.null_embedding), C.ptr(out), B, Lq, D, EOS_ID,
                   int(self.use_fixed_special), 0, s)
        elif save_for_backward:
            raise NotImplementedError("training runs through TextBoostModel (null-embedding override on)")
        if save_for_backward:
            self._ctx.append((ids, saved, x, stf))
        return out.view(B, Lq, D)

    @staticmethod
    def split_ctx(ctx, n_first: int):
        """Split the saved context of ONE forward over a concatenated batch [first n_first prompts | rest] into two
        contexts that backward() accepts independently (every saved tensor is row-major over prompts x tokens, so the
        halves are views).  The trainer encodes the instance and the prior prompts in one pass -- the kernels at 616 or
        1232 rows cost the same latency -- and runs their backwards at different times."""
        ids, saved, xf, stf = ctx
        B, Lq = ids.shape
        m = n_first * Lq

        def cut(t, lo):
            if t is None:
                return None
            if t.dim() == 3 and t.shape[0] == B:      # lse [B, heads, L]
                return t[:n_first] if lo else t[n_first:]
            return t[:m] if lo else t[m:]             # [B*L, ...] activations / statistics

        halves = []
        for lo in (True, False):
            halves.append((ids[:n_first] if lo else ids[n_first:], [tuple(cut(t, lo) for t in layer) for layer in saved],
                           cut(xf, lo), cut(stf, lo)))
        return halves

    def pop_ctx(self):
        """Detach the context of the most recent forward(save_for_backward=True) (autograd wrappers keep it
        on their own ctx so several forwards can be outstanding, as in train_textboost.py:1054-1100)."""
        return self._ctx.pop()

    # ------------------------------------------------------------------ backward
    def backward(self, d_out: torch.Tensor, ctx=None):
        """d_out fp32 [B, L, D] (consumed / overwritten).  Accumulates into state.grads."""
        if ctx is None:
            assert self._ctx, "forward(save_for_backward=True) must precede backward"
            ctx = self._ctx.pop()
        ids, saved, xf, stf = ctx
        B, Lq = ids.shape
        D, M = self.D, B * Lq
        st = self.state
        s = C.stream_ptr()
        d_out = d_out.contiguous().view(M, D)
        # The LoRA weight-gradient kernel of a layer hangs off the chain (nothing downstream reads dA / dB): it runs on
        # a helper stream next to the following layer's GEMMs instead of adding its ~10 us to each of the 12 links of
        # the latency-bound tail.  Its operands are kept alive until the join below.
        main = torch.cuda.current_stream() if d_out.is_cuda else None
        helper = self._helper_stream(main) if main is not None and self.r else None
        keep = []
        C.call("tb_null_override", C.ptr(ids), None, C.ptr(d_out), B, Lq, D, EOS_ID,
               int(self.use_fixed_special), 1, s)
        # every LayerNorm backward also writes the fp16 copy of its result: the next dgrad GEMM's operand
        g16 = torch.empty((M, D), device=self.device, dtype=POLICY.act)
        g = ops.layernorm_bwd_clip(d_out, xf, self.lnf[0], stf, out16=g16)
        for l in reversed(range(self.nl)):
            L = self.layers[l]
            x, st1, y_ext, qkv, x2, st2, u, o, lse = saved[l]
            da = ops.gemm(g16, L["wf2_t"])
            du = torch.empty_like(da)
            C.call("tb_act_bwd_f16", C.ptr(u), C.ptr(da), C.ptr(du), u.numel(), self.act, s)
            dy2 = ops.gemm(du, L["wf1_t"])
            g16 = torch.empty((M, D), device=self.device, dtype=POLICY.act)
            g = ops.layernorm_bwd_clip(dy2, x2, L["ln2"][0], st2, add=g, out=g, out16=g16)
            do = ops.gemm(g16, L["wo_t"])  # [M, Ko]: d(o) and, in the extension columns, d(o A_o^T)
            if self.has_o:
                ro = self.Tq * self.r
                C.call("tb_lora_grad", C.ptr(g16), C.ptr(o), C.ptr(do), self.Ko, C.ptr(st.B(l, st.grads)[self.Tq]),
                       C.ptr(st.A(l, st.grads)[ro:]), M, 1, 1, D, self.r, self.scaling, s)
                C.call("tb_lora_dx", C.ptr(do), self.Ko, C.ptr(st.A(l)[ro:]), M, D, self.r, s)
            dqkv = torch.empty_like(qkv)
            q3, d3 = qkv.view(B, Lq, 3 * D), dqkv.view(B, Lq, 3 * D)
            # 77 tokens = one KV tile: dQ is written once, as fp16, straight into the fused gradient tensor
            ops.attn_bwd(q3[..., :D], q3[..., D:2 * D], q3[..., 2 * D:], o.view(B, Lq, self.Ko)[..., :D],
                         do.view(B, Lq, self.Ko)[..., :D], lse, self.heads, dk=d3[..., D:2 * D],
                         dv=d3[..., 2 * D:], causal=True, dq_out=d3[..., :D])
            dy_ext = ops.gemm(dqkv, L["wqkv_t"])
            g16 = torch.empty((M, D), device=self.device, dtype=POLICY.act) if l else None
            if self.Tq:
                if helper is not None:
                    ready = torch.cuda.Event()
                    ready.record(main)
                    keep.append((dqkv, y_ext, dy_ext))
                    with torch.cuda.stream(helper):
                        helper.wait_event(ready)
                        C.call("tb_lora_grad", C.ptr(dqkv), C.ptr(y_ext), C.ptr(dy_ext), self.Kext,
                               C.ptr(st.B(l, st.grads)), C.ptr(st.A(l, st.grads)), M, 3, self.qkv_mask, D, self.r,
                               self.scaling, C.stream_ptr())
                else:
                    C.call("tb_lora_grad", C.ptr(dqkv), C.ptr(y_ext), C.ptr(dy_ext), self.Kext,
                           C.ptr(st.B(l, st.grads)), C.ptr(st.A(l, st.grads)), M, 3, self.qkv_mask, D, self.r,
                           self.scaling, s)
                # the down-projection's input-gradient (dy += dxa A) rides on the LayerNorm backward
                g = ops.layernorm_bwd_clip(dy_ext, x, L["ln1"][0], st1, add=g, out=g, out16=g16,
                                           lora_a=st.A(l)[:self.Tq * self.r])
            else:
                g = ops.layernorm_bwd_clip(dy_ext[:, :D], x, L["ln1"][0], st1, add=g, out=g, out16=g16)
            saved[l] = None
        if st.n_rows:
            C.call("tb_clip_embed_grad", C.ptr(ids), C.ptr(g), C.ptr(st.rows(st.grads)), M, D, self.n_base, s)
        if helper is not None:
            main.wait_stream(helper)
        keep.clear()
        return g.view(B, Lq, D)

    def _helper_stream(self, main):
        """One helper stream per stream backward() is called on (the prior-prompt backward runs on the trainer's side
        stream while the UNet owns the main one)."""
        if not hasattr(self, "_helpers"):
            self._helpers = {}
        key = main.cuda_stream
        if key not in self._helpers:
            self._helpers[key] = torch.cuda.Stream(device=self.device)
        return self._helpers[key]
