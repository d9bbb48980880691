"""Multi-GPU host logic: one process per GPU, batch-sharded snippets, and the path's only exchange --
the 4-float loss vector -- all-reduced over NCCL (NVLink/NVSwitch).  Mirrors what the reference gets
from tf.distribute.MirroredStrategy (model/model_util/distributer.py:5-110): contiguous per-replica
batch shards, every replica normalising by the GLOBAL batch (losses.py:49), replica losses summed
(distributer.py:93-96).  Per-snippet outputs (synth_ms, d_depth_ms, d_disp_ms, d_pose) stay local:
they feed each rank's own net backward."""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `global_batch` snippets for `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(features: Dict, predictions: Dict, rank: int, world: int) -> Tuple[Dict, Dict]:
    """Slice a global batch of features/predictions down to this rank's snippets (views, no copy)."""
    B = features["image5d"].shape[0]
    lo, hi = shard_bounds(B, rank, world)
    f = {k: v[lo:hi] for k, v in features.items()}
    p = {}
    for k, v in predictions.items():
        p[k] = [t[lo:hi] for t in v] if isinstance(v, (list, tuple)) else v[lo:hi]
    return f, p


def allreduce_losses(loss_vec: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the per-rank loss vector [total, L1, SSIM, smoothe] in place.  Each rank's entries are already
    divided by the GLOBAL batch, so the sum over ranks is the global-batch mean the reference reports."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(loss_vec, op=dist.ReduceOp.SUM, group=group)
    return loss_vec


def allreduce_gradient_buckets(buckets: Sequence[torch.Tensor], group=None, async_op: bool = True):
    """Sum net-gradient buckets across ranks (what MirroredStrategy does inside apply_gradients,
    train_val.py:86).  The nets are outside this path; provided so a caller can overlap it with the
    next fused launch.  Returns the work handles (empty when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return []
    return [dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group, async_op=async_op) for b in buckets]
