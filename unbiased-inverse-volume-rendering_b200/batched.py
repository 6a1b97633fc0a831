"""Ray-batch rendering -- the rank-2 "next" row of SURVEY §8(f): host mirror of python/batched.py.

Reference                                                     -> here
  render_batch(batch_size, scene, sensors, film_size, params, integrator, ..., seed, seed_grad,
               spp, spp_grad)                (batched.py:88-131; called at optimize.py:334-340)
                                                                -> render_batch
  sample_batch_pixels                       (batched.py:397-423) -> sample_batch_pixels (host restatement of
                                                                    the index sampler; the kernels redo it on chip)
  _BatchedRenderOp.eval / .backward         (batched.py:13-85)   -> _BatchedRenderOp (torch.autograd.Function)
  gather_ref_values                         (optimize.py:90-107) -> gather_ref_values

The wavefront is a batch of B (sensor, pixel) pairs drawn uniformly, spp samples each, film
(B x 1) with a box filter.  Everything per sample -- which sensor / pixel, the sub-pixel offsets
(decorrelated between primal and adjoint), the path itself -- is regenerated inside the CUDA
kernels from counter-based streams (include/uivr.h: uivr_batch_desc); nothing is staged per ray.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native
from .integrator import (ALBEDO_SUFFIX, SIGMA_T_SUFFIX, NeRFIntegrator, Scene, VolpathSimpleIntegrator, _find_key,
                         _stream)
from .scene import Sensor

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_PCG_MULT = np.uint64(0x5851F42D4C957F2D)


# ---- host restatement of the `independent` sampler (PCG32 seeded through TEA), vectorised ----

def _tea(v0: np.ndarray, v1: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """mi.sample_tea_32 (4 rounds), uint32 arrays."""
    v0 = v0.astype(np.uint32).copy()
    v1 = v1.astype(np.uint32).copy()
    s = np.uint32(0)
    with np.errstate(over="ignore"):
        for _ in range(4):
            s = np.uint32((int(s) + 0x9E3779B9) & 0xFFFFFFFF)
            v0 += ((v1 << np.uint32(4)) + np.uint32(0xA341316C)) ^ (v1 + s) ^ ((v1 >> np.uint32(5)) + np.uint32(0xC8013EA4))
            v1 += ((v0 << np.uint32(4)) + np.uint32(0xAD90777D)) ^ (v0 + s) ^ ((v0 >> np.uint32(5)) + np.uint32(0x7E95761E))
    return v0, v1


class _Pcg32:
    """PCG32 streams, one per array element (sampler.seed(seed, wavefront_size))."""

    def __init__(self, seed: int, n: int):
        v0, v1 = _tea(np.full(n, seed & 0xFFFFFFFF, dtype=np.uint32), np.arange(n, dtype=np.uint32))
        self.inc = (v1.astype(np.uint64) << np.uint64(1)) | np.uint64(1)
        self.state = np.zeros(n, dtype=np.uint64)
        self._next()
        with np.errstate(over="ignore"):
            self.state += v0.astype(np.uint64)
        self._next()

    def _next(self) -> np.ndarray:
        old = self.state
        with np.errstate(over="ignore"):
            self.state = old * _PCG_MULT + self.inc
        xs = (((old >> np.uint64(18)) ^ old) >> np.uint64(27)).astype(np.uint32)
        rot = (old >> np.uint64(59)).astype(np.uint32)
        return (xs >> rot) | (xs << ((np.uint32(0) - rot) & np.uint32(31)))

    def next_1d(self) -> np.ndarray:
        return ((self._next() >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)


def sample_batch_pixels(batch_size: int, n_sensors: int, film_size: Tuple[int, int], seed: int):
    """batched.py:397-423: (sensor_idx [B], pixels [B, 2] as (x, y)) of the batch drawn for `seed`."""
    w, h = int(film_size[0]), int(film_size[1])
    s = _Pcg32(_native.tea32(seed, 5), batch_size)  # sub_seed_0 = tea32(seed, 17*0 + 5)
    u0, u1, u2 = s.next_1d(), s.next_1d(), s.next_1d()
    sensor_idx = np.minimum((np.float32(n_sensors) * u0).astype(np.uint32), np.uint32(n_sensors - 1))
    px = np.minimum((np.float32(w) * u1).astype(np.uint32), np.uint32(w - 1))
    py = np.minimum((np.float32(h) * u2).astype(np.uint32), np.uint32(h - 1))
    return sensor_idx, np.stack([px, py], axis=1)


def sensor_table(sensors: Sequence[Sensor]) -> np.ndarray:
    """[n, 16] float32 rows origin[3] left[3] up[3] dir[3] tan_x tan_y near_clip 0 (uivr_batch_desc.sensors)."""
    w, h = sensors[0].width, sensors[0].height
    rows = []
    for s in sensors:
        if (s.width, s.height) != (w, h):
            raise ValueError("all sensors of a ray batch must share the film size (batched.py:428)")
        f = s.frame()
        rows.append(np.concatenate([f["cam_origin"], f["cam_left"], f["cam_up"], f["cam_dir"],
                                    [f["tan_x"], f["tan_y"], f["near_clip"], 0.0]]).astype(np.float32))
    return np.stack(rows)


def gather_ref_values(ref_images: torch.Tensor, sensor_idx, pixel_idx) -> torch.Tensor:
    """optimize.py:90-107: reference colours [B, C] of the batch from ref_images [n, H, W, C]."""
    assert ref_images.dim() == 4 and ref_images.shape[-1] in (3, 4)
    si = torch.as_tensor(np.asarray(sensor_idx).astype(np.int64), device=ref_images.device)
    p = torch.as_tensor(np.asarray(pixel_idx).astype(np.int64), device=ref_images.device)
    return ref_images[si, p[:, 1], p[:, 0]]


class _BatchedRenderOp(torch.autograd.Function):
    """batched.py:13-85: eval = primal batch render (detached); backward = render_batch_backward on a
    decorrelated set of rays through the same pixels."""

    @staticmethod
    def forward(ctx, sigma_t, albedo, scene, integrator, batch, seed, seed_grad, spp, spp_grad, keys, shard, reducer):
        params = {keys[0]: sigma_t, keys[1]: albedo}
        image = _launch(scene, integrator, params, batch, seed, spp, None, None, shard)
        ctx.save_for_backward(sigma_t, albedo)
        ctx.meta = (scene, integrator, batch, seed_grad, spp_grad, keys, shard, reducer)
        return image

    @staticmethod
    def backward(ctx, grad_image):
        sigma_t, albedo = ctx.saved_tensors
        scene, integrator, batch, seed_grad, spp_grad, keys, shard, reducer = ctx.meta
        params = {keys[0]: sigma_t, keys[1]: albedo}
        dsig, dalb = _launch(scene, integrator, params, batch, seed_grad, spp_grad, grad_image, None, shard)
        if reducer is not None:
            dsig, dalb = reducer(dsig, dalb)
        return (dsig, dalb) + (None,) * 10


def _launch(scene: Scene, integrator: VolpathSimpleIntegrator, params, batch, seed, spp, grad_image, sample_out,
            shard=None):
    table, film_size, batch_size, batch_seed = batch
    nerf = isinstance(integrator, NeRFIntegrator)  # render_batch takes any registered integrator (batched.py:110-112)
    sig, alb = scene.check_params(params, integrator.second_suffix)
    props = integrator.props()
    scene.bind(None, None if nerf else props)
    scene.update_medium(sig.detach())
    scene.ctx.set_batch(table, film_size[0], film_size[1], batch_size, batch_seed)
    try:
        sp = None if sample_out is None else sample_out.data_ptr()
        if grad_image is None:
            image = torch.empty((batch_size, 3), dtype=torch.float32, device=sig.device)
            if nerf:
                scene.ctx.nerf_forward(props, alb.detach().data_ptr(), seed, spp, image.data_ptr(), sp, shard, _stream())
            else:
                scene.ctx.render_forward(alb.detach().data_ptr(), seed, spp, image.data_ptr(), sp, shard, _stream())
            return image
        g = grad_image.to(dtype=torch.float32).contiguous()
        if tuple(g.shape) != (batch_size, 3):
            raise ValueError(f"grad_in must have shape {(batch_size, 3)}")
        dsig, dalb = torch.empty_like(sig), torch.empty_like(alb)
        if nerf:
            scene.ctx.nerf_backward(props, alb.detach().data_ptr(), g.data_ptr(), seed, spp, dsig.data_ptr(),
                                    dalb.data_ptr(), sp, shard, _stream())
        else:
            scene.ctx.render_backward(alb.detach().data_ptr(), g.data_ptr(), seed, spp, dsig.data_ptr(),
                                      dalb.data_ptr(), sp, shard, _stream())
        return dsig, dalb
    finally:
        scene.ctx.set_batch(None)


def render_batch(batch_size: int, scene: Scene, sensors: Sequence[Sensor], params: Dict[str, torch.Tensor],
                 integrator: VolpathSimpleIntegrator, seed: int = 0, seed_grad: int = 0, spp: int = 0,
                 spp_grad: int = 0, shard=None, reducer=None):
    """batched.py:88-131.  Returns (image [B, 3] differentiable w.r.t. the two grids, sensor_idx [B],
    pixels [B, 2]) -- the reference returns the same triple next to its film / sampler objects
    (batched.py:53-56), the indices feeding gather_ref_values (optimize.py:341).

    Data-parallel ray batches (the reference's production mode on several GPUs): every rank calls this with
    the SAME seed -- hence the same (sensor, pixel) batch and the same counter-based RNG streams -- and its own
    `shard` = sharding.pixel_shard(rank, world): batch element b is rendered by rank (b // block) % world, the
    other rows of this rank's image stay 0 (a per-element loss needs no forward collective; `gather_image` sums
    the partial images when the whole batch is wanted).  `reducer(dsig, dalb)` (e.g. sharding.all_reduce_pair)
    runs on the parameter gradients in the backward pass."""
    if spp <= 0:
        raise ValueError("spp must be positive")
    if spp_grad == 0:
        spp_grad = spp
    if seed_grad == 0:
        seed_grad = _native.tea32(seed, 1)  # de-correlate the primal and differential phase (batched.py:119-121)
    elif seed_grad == seed:
        raise Exception("The primal and differential seed should be different "
                        "to ensure unbiased gradient computation!")  # batched.py:122-124
    film_size = (sensors[0].width, sensors[0].height)
    table = sensor_table(sensors)
    sensor_idx, pixels = sample_batch_pixels(batch_size, len(sensors), film_size, seed)
    k_sig, k_alb = _find_key(params, SIGMA_T_SUFFIX), _find_key(params, integrator.second_suffix)
    batch = (table, film_size, int(batch_size), seed & 0xFFFFFFFF)
    image = _BatchedRenderOp.apply(params[k_sig], params[k_alb], scene, integrator, batch, seed, seed_grad,
                                   spp, spp_grad, (k_sig, k_alb), shard, reducer)
    return image, sensor_idx, pixels
