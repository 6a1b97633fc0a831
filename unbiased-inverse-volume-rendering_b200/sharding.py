"""Multi-GPU plumbing for the volpathsimple path: one process per GPU, pixels sharded across
ranks, parameter gradients summed with ONE all-reduce (NCCL over NVLink on the GPU box, gloo in
the CPU tests).

The reference is single-process / single-GPU; what shards naturally is the
wavefront of `mi.render` (one independent Monte-Carlo sample per lane, batched.py:378-393).
RNG streams are keyed by the GLOBAL sample index (sampler.seed(seed, wavefront_size)), so the
union of the shards reproduces the single-GPU samples exactly; only the order of the
floating-point gradient sums changes.

 - forward : every rank renders its pixels; pixels of other ranks stay 0 in its image.  For a
             per-pixel separable loss (losses.py:7-11) no forward collective is needed: the
             loss gradient of a pixel depends on that pixel only.  `gather_image` sums the
             disjoint partial images when the full picture is wanted.
 - backward: every rank scatters the gradients of its samples into its own full-size
             [d sigma_t | d albedo] buffer; `GradientBuffer.all_reduce` sums the buffers.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

# Interleaving granularity (pixels).  Path cost varies strongly across the image (rays that miss
# the medium are ~free), so contiguous row blocks are badly balanced, and a block size that divides
# the row length hands whole COLUMN strips to one rank (64-pixel blocks of a 512-wide film on 8
# ranks: rank 0 renders only the empty left border).  Pure pixel interleaving (block 1) gives every
# rank the same statistical mix and -- measured on config 3, scripts/shard_bench.py -- the same
# kernel times as the unsharded render (block 64: +50 % forward, +13 % backward at 4 ranks).
DEFAULT_BLOCK = 1


def pixel_shard(rank: int, world_size: int, block: int = DEFAULT_BLOCK) -> Optional[Tuple[int, int, int]]:
    """uivr_shard for this rank: pixel p belongs to rank (p // block) % world_size."""
    if world_size <= 1:
        return None
    if not (0 <= rank < world_size) or block < 1:
        raise ValueError("need 0 <= rank < world_size and block >= 1")
    return (int(rank), int(world_size), int(block))


def owned_pixel_mask(n_pixels: int, shard: Optional[Tuple[int, int, int]], device=None) -> torch.Tensor:
    """Boolean mask [n_pixels] of the pixels a shard renders (host-side mirror of the kernels'
    slot_to_pixel)."""
    p = torch.arange(n_pixels, device=device)
    if shard is None:
        return torch.ones(n_pixels, dtype=torch.bool, device=device)
    rank, count, block = shard
    return (p // block) % count == rank


class GradientBuffer:
    """[d sigma_t (Z,Y,X,1) | d albedo (Z,Y,X,3)] in one flat fp32 allocation, so that the
    gradient exchange is a single collective over 4*Z*Y*X floats."""

    def __init__(self, res: Tuple[int, int, int], device):
        x, y, z = (int(r) for r in res)
        n = x * y * z
        self.flat = torch.zeros(4 * n, dtype=torch.float32, device=device)
        self.dsigma = self.flat[:n].view(z, y, x, 1)
        self.dalbedo = self.flat[n:].view(z, y, x, 3)

    def views(self):
        return self.dsigma, self.dalbedo

    def all_reduce(self, group=None, async_op: bool = False):
        """Sum over ranks (ncclAllReduce on the GPU box).  No-op without a process group."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return None


def gather_image(image: torch.Tensor, group=None) -> torch.Tensor:
    """Full image from the ranks' disjoint partial images (pixels of other ranks are 0)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(image, op=dist.ReduceOp.SUM, group=group)
    return image


def all_reduce_pair(dsigma: torch.Tensor, dalbedo: torch.Tensor, group=None):
    """`reducer` for render(..., reducer=) / render_batch(..., reducer=): sum the two parameter gradients over
    the ranks (two collectives; GradientBuffer packs them into one for the bench loop)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(dsigma, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(dalbedo, op=dist.ReduceOp.SUM, group=group)
    return dsigma, dalbedo
