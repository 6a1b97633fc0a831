"""Optimisation step around the render path -- the rank-1 "next" row of SURVEY §8(f).

Reference                                               -> here
  mi.ad.Adam(lr, params) / opt.set_learning_rate / opt.step()
        (opt_config.py:46-48, optimize.py:313, :329, :352)  -> Adam (fused CUDA kernel, uivr_adam_step)
  enforce_valid_params (optimize.py:169-179, :353)           -> fused into the same kernel (clip range)
  OptimizationConfig.learning_rates (opt_config.py:50-69)    -> learning_rates
  losses.l1 (losses.py:7-8)                                  -> l1_loss_grad
  the loop body of run_optimization (optimize.py:325-354)    -> optimization_step

One step = for every given sensor: render at `seed`, L1 against its reference image, backward at
`seed_grad` (seeds per optimize.py:327-328), gradients summed over the views; then one Adam
update per parameter tensor with its own learning rate, the projection to the legal range and
the rebuild of the medium's lookup structures (params.update(), optimize.py:354).  The reference
itself renders ONE random sensor (or one ray batch) per iteration; several views per step is
BASELINE.json's config 4.  PyTorch only holds the tensors; the update runs in libuivr.so.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _native
from .integrator import ALBEDO_SUFFIX, SIGMA_T_SUFFIX, Scene, VolpathSimpleIntegrator, _stream
from .scene import Sensor

LAST25_STEPS = (0.75, 0.85, 0.95)  # opt_config.py:54-55


def learning_rates(lr: float, param_keys: Sequence[str], it_i: int, n_iter: int, schedule: Optional[str] = None,
                   param_lr_factors: Optional[Dict[str, float]] = None) -> Dict[str, float]:
    """OptimizationConfig.learning_rates (opt_config.py:50-69): per-key rate = schedule factor x
    per-key factor (scene_config.param_lr_factors) x base rate.  schedule: None | 'constant' | 'last25'."""
    factor = 1.0
    if schedule not in (None, "constant"):
        if schedule != "last25":
            raise ValueError(f"Unsupported schedule: {schedule}")
        t = it_i / (n_iter - 1)
        for s in LAST25_STEPS:
            if t >= s:
                factor *= 0.5
    f = param_lr_factors or {}
    return {k: factor * f.get(k, 1.0) * lr for k in param_keys}


def param_bounds(key: str, max_density: float = 250.0) -> Tuple[float, float]:
    """enforce_valid_params (optimize.py:169-179): legal range of a parameter tensor."""
    if key.endswith(SIGMA_T_SUFFIX):
        return 0.0, float(max_density)
    if key.endswith("emission.data"):
        return 0.0, float("inf")
    if key.endswith(ALBEDO_SUFFIX):
        return 0.0, 1.0
    raise ValueError(key)


class Adam:
    """mi.ad.Adam as the reference uses it: constructed from (lr, params), per-key rates through
    set_learning_rate, step() consumes the gradients.  State (m, v, t) lives on the device."""

    def __init__(self, lr: float, params: Dict[str, torch.Tensor], beta_1: float = 0.9, beta_2: float = 0.999,
                 epsilon: float = 1e-8):
        self.params = params
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.lr = {k: float(lr) for k in params}
        self.t = {k: 0 for k in params}   # step counter per parameter (mi.ad.Adam keeps its state per key)
        self.m = {k: torch.zeros_like(p) for k, p in params.items()}
        self.v = {k: torch.zeros_like(p) for k, p in params.items()}
        for k, p in params.items():
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise ValueError(f"{k} must be a contiguous float32 CUDA tensor")

    def set_learning_rate(self, lr):
        if isinstance(lr, dict):
            for k, v in lr.items():
                if k not in self.lr:
                    raise KeyError(k)
                self.lr[k] = float(v)
        else:
            self.lr = {k: float(lr) for k in self.lr}

    def items(self):
        return self.params.items()

    def replace(self, key: str, value: torch.Tensor):
        """opt[key] = value with a new shape (multires upsampling, optimize.py:239-242): the moments and
        the step counter of that parameter start over, as mi.ad.Optimizer resets a re-sized parameter."""
        if key not in self.params:
            raise KeyError(key)
        value = value.detach().to(torch.float32).contiguous()
        self.params[key] = value
        self.m[key] = torch.zeros_like(value)
        self.v[key] = torch.zeros_like(value)
        self.t[key] = 0

    def step(self, ctx: _native.Context, grads: Dict[str, torch.Tensor], max_density: float = 250.0):
        """opt.step() + enforce_valid_params in one pass per tensor."""
        for k, p in self.params.items():
            self.t[k] += 1
            g = grads[k]
            if g.shape != p.shape or g.dtype != torch.float32 or not g.is_contiguous():
                raise ValueError(f"gradient of {k} must match its parameter")
            lo, hi = param_bounds(k, max_density)
            ctx.adam_step(p.data_ptr(), g.data_ptr(), self.m[k].data_ptr(), self.v[k].data_ptr(), p.numel(),
                          self.lr[k], self.beta_1, self.beta_2, self.epsilon, self.t[k], lo,
                          hi if hi != float("inf") else 3.4028234663852886e38, _stream())


def l1_loss_grad(image: torch.Tensor, ref: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """losses.l1 (losses.py:7-8): mean |image - ref| and its gradient w.r.t. the image."""
    d = image - ref
    return d.abs().mean(), torch.sign(d) / d.numel()


def optimization_step(scene: Scene, integrator: VolpathSimpleIntegrator, opt: Adam, sensors: Sequence[Sensor],
                      refs: Sequence[torch.Tensor], it_i: int, spp: int, spp_grad: int = 0, base_seed: int = 1234,
                      max_density: float = 250.0, grads: Optional[Dict[str, torch.Tensor]] = None) -> float:
    """Loop body of run_optimization (optimize.py:325-354) over the given views.  Returns the mean loss."""
    params = opt.params
    k_sig = next(k for k in params if k.endswith(SIGMA_T_SUFFIX))
    k_alb = next(k for k in params if k.endswith(integrator.second_suffix))  # albedo, or emission for `nerf`
    spp_grad = spp_grad or spp
    if grads is None:
        grads = {k: torch.zeros_like(p) for k, p in params.items()}
    else:
        for g in grads.values():
            g.zero_()
    view = {k_sig: torch.empty_like(params[k_sig]), k_alb: torch.empty_like(params[k_alb])}
    loss_sum = torch.zeros((), device=params[k_sig].device)
    for j, (sensor, ref) in enumerate(zip(sensors, refs)):
        n = it_i * len(sensors) + j
        seed, seed_grad = _native.tea32(2 * n, base_seed), _native.tea32(2 * n + 1, base_seed)  # optimize.py:327-328
        image = integrator.render(scene, params, sensor=sensor, seed=seed, spp=spp)
        loss, g_img = l1_loss_grad(image, ref)
        integrator.render_backward(scene, params, g_img, sensor=sensor, seed=seed_grad, spp=spp_grad,
                                   out=(view[k_sig], view[k_alb]))
        grads[k_sig] += view[k_sig]
        grads[k_alb] += view[k_alb]
        loss_sum += loss
    opt.step(scene.ctx, grads, max_density)                 # optimize.py:352-353
    scene.update_medium(params[k_sig], force=True)          # params.update(), optimize.py:354
    return float(loss_sum.item()) / max(1, len(sensors))
