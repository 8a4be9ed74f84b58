"""Data parallelism of the TextBoost step: one process per GPU, frozen weights replicated, the batch
sharded by rows, ONE all-reduce per step of the flat [LoRA | added embedding rows] gradient buffer.

Replaces accelerate's DDP wrap of the text encoder (/root/reference/train_textboost.py:919-922), which
all-reduces every requires_grad parameter — the whole 49,408 x 768 embedding matrix, 152.7 MB — and then
zeroes the frozen rows (:1109-1117).  Reducing only the rows that survive that mask gives the
same parameters after the step and moves ~0.9 MB.  The SUM is taken here; the 1/world of DDP's average and
the GradScaler unscale are applied by the fused AdamW kernel that consumes the buffer
(tb_adamw_fused_step(inv_world=...)), so there is no separate scaling pass.

Rank r of W owns rows [r*b, (r+1)*b) of a global batch of W*b (the reference shards its dataset the same
way, textboost/dataset.py:846-870: indices[rank::world] of an equal-length stream).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None, device: Optional[torch.device] = None):
    """torchrun-style bootstrap (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT).
    Returns (rank, world, local_rank); a single-process run needs no process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = device if device is not None else torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, world, local_rank


class GradSync:
    """The step's single exchange: sum the flat gradient buffer over the data-parallel group."""

    def __init__(self, process_group=None):
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0

    @property
    def inv_world(self) -> float:
        return 1.0 / self.world

    def all_reduce_(self, flat_grads: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=self.pg)
        return flat_grads

    def shard_rows(self, global_batch: int):
        """slice of the global batch this rank trains on."""
        assert global_batch % self.world == 0, "the global batch must split evenly (DDP AVG == global mean)"
        b = global_batch // self.world
        return slice(self.rank * b, (self.rank + 1) * b)

    def payload_bytes(self, flat_grads: torch.Tensor) -> int:
        return flat_grads.numel() * flat_grads.element_size()
