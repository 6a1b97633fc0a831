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

import math
import os
from typing import Callable, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native
from .exr import read_exr, write_exr
from .integrator import ALBEDO_SUFFIX, SIGMA_T_SUFFIX, Scene, VolpathSimpleIntegrator, _stream, load_dict, render
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
            g = grads.get(k)
            if g is None:        # mi.ad.Adam.step skips a parameter whose gradient is empty (not touched by the render)
                continue
            self.t[k] += 1
            if g.shape != p.shape or g.dtype != torch.float32 or not g.is_contiguous():
                raise ValueError(f"gradient of {k} must match its parameter")
            lo, hi = param_bounds(k, max_density)
            ctx.adam_step(p.data_ptr(), g.data_ptr(), self.m[k].data_ptr(), self.v[k].data_ptr(), p.numel(),
                          self.lr[k], self.beta_1, self.beta_2, self.epsilon, self.t[k], lo,
                          hi if hi != float("inf") else 3.4028234663852886e38, _stream())
            # the kernel wrote through the raw pointer: tell torch (and Scene.update_medium's cache key)
            torch.autograd.graph.increment_version(p)


class SGD:
    """mi.ad.SGD as `OptimizationConfig.optimizer` can construct it (opt_config.py:46-48; the reference's
    runs all use Adam): value -= lr * (momentum-filtered) gradient, then the projection of
    enforce_valid_params.  Plain torch element-wise updates -- SGD is not part of any measured
    configuration and has no kernel of its own."""

    def __init__(self, lr: float, params: Dict[str, torch.Tensor], momentum: float = 0.0):
        self.params = params
        self.momentum = float(momentum)
        self.lr = {k: float(lr) for k in params}
        self.state = {}

    set_learning_rate = Adam.set_learning_rate
    items = Adam.items

    def replace(self, key: str, value: torch.Tensor):
        if key not in self.params:
            raise KeyError(key)
        self.params[key] = value.detach().to(torch.float32).contiguous()
        self.state.pop(key, None)

    def step(self, ctx, grads: Dict[str, torch.Tensor], max_density: float = 250.0):
        for k, p in self.params.items():
            g = grads.get(k)
            if g is None:
                continue
            if self.momentum != 0.0:
                g = self.state[k] = g.clone() if k not in self.state else self.state[k].mul_(self.momentum).add_(g)
            lo, hi = param_bounds(k, max_density)
            p.detach().add_(g, alpha=-self.lr[k]).clamp_(lo, hi)


def l1_loss_grad(image: torch.Tensor, ref: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """losses.l1 (losses.py:7-8): mean |image - ref| and its gradient w.r.t. the image."""
    d = image - ref
    return d.abs().mean(), torch.sign(d) / d.numel()


def optimization_step(scene: Scene, integrator: VolpathSimpleIntegrator, opt: Adam, sensors: Sequence[Sensor],
                      refs: Sequence[torch.Tensor], it_i: int, spp: int, spp_grad: int = 0, base_seed: int = 1234,
                      max_density: float = 250.0, grads: Optional[Dict[str, torch.Tensor]] = None,
                      lanes: Optional[Sequence["ViewLane"]] = None) -> float:
    """Loop body of run_optimization (optimize.py:325-354) over the given views.  Returns the mean loss.

    `lanes` (make_view_lanes): the views alternate between several contexts of the same medium, each on a CUDA stream
    of its own.  The views of an iteration are independent until their gradients are summed, and a persistent launch
    spends 0.6-0.8 ms draining its pool after its work queue is empty: with two lanes the
    kernels of the next view start on the SMs the previous view's launch has already left.  Same seeds, same samples;
    only the order of the float sums of the gradients changes."""
    params = opt.params
    k_sig = next(k for k in params if k.endswith(SIGMA_T_SUFFIX))
    k_alb = next(k for k in params if k.endswith(integrator.second_suffix))  # albedo, or emission for `nerf`
    spp_grad = spp_grad or spp
    if grads is None:
        grads = {k: torch.zeros_like(p) for k, p in params.items()}
    else:
        for g in grads.values():
            g.zero_()
    if lanes:
        return _optimization_step_lanes(scene, integrator, opt, sensors, refs, it_i, spp, spp_grad, base_seed, max_density,
                                        grads, lanes, k_sig, k_alb)
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


class ViewLane:
    """One of several contexts that render the views of an iteration side by side (optimization_step(lanes=...))."""

    def __init__(self, scene: Scene, params: Dict[str, torch.Tensor]):
        self.scene = scene
        self.stream = torch.cuda.Stream(device=scene.device)
        self.view = {k: torch.empty_like(p) for k, p in params.items()}
        self.acc = {k: torch.zeros_like(p) for k, p in params.items()}
        self.loss = torch.zeros((), device=next(iter(params.values())).device)


def make_view_lanes(scene: Scene, params: Dict[str, torch.Tensor], count: int = 2) -> list:
    """`count` lanes for optimization_step: the given scene + (count - 1) more contexts of the same volume.  Each
    context owns its copy of the lookup structures and scratch (about 4 GB at 256^3)."""
    return [ViewLane(scene if i == 0 else Scene(scene.volume, scene.device), params) for i in range(count)]


def _optimization_step_lanes(scene, integrator, opt, sensors, refs, it_i, spp, spp_grad, base_seed, max_density, grads,
                             lanes, k_sig, k_alb) -> float:
    params = opt.params
    main = torch.cuda.current_stream()
    for lane in lanes:
        lane.stream.wait_stream(main)       # parameters (and the previous iteration's update) are ready
        with torch.cuda.stream(lane.stream):
            for a in lane.acc.values():
                a.zero_()
            lane.loss.zero_()
    for j, (sensor, ref) in enumerate(zip(sensors, refs)):
        lane = lanes[j % len(lanes)]
        n = it_i * len(sensors) + j
        seed, seed_grad = _native.tea32(2 * n, base_seed), _native.tea32(2 * n + 1, base_seed)  # optimize.py:327-328
        with torch.cuda.stream(lane.stream):
            image = integrator.render(lane.scene, params, sensor=sensor, seed=seed, spp=spp)
            loss, g_img = l1_loss_grad(image, ref)
            integrator.render_backward(lane.scene, params, g_img, sensor=sensor, seed=seed_grad, spp=spp_grad,
                                       out=(lane.view[k_sig], lane.view[k_alb]))
            lane.acc[k_sig] += lane.view[k_sig]
            lane.acc[k_alb] += lane.view[k_alb]
            lane.loss += loss
    loss_sum = torch.zeros((), device=params[k_sig].device)
    for lane in lanes:
        main.wait_stream(lane.stream)
        grads[k_sig] += lane.acc[k_sig]
        grads[k_alb] += lane.acc[k_alb]
        loss_sum += lane.loss
    opt.step(scene.ctx, grads, max_density)                 # optimize.py:352-353
    for lane in lanes:                                       # params.update(), optimize.py:354 (every lane's copy)
        if lane.scene._scene_key is not None:                # (a lane that has not rendered yet builds it at its first view)
            lane.scene.update_medium(params[k_sig], force=True)
    return float(loss_sum.item()) / max(1, len(sensors))


# ======================================================================================
# The run around the step: python/optimize.py:14-88, :110-166, :255-365
# ======================================================================================

class PCG32:
    """mi.scalar_rgb.PCG32 as optimize.py:291, :343 uses it to pick the sensor of an iteration."""
    _MULT, _M64 = 0x5851F42D4C957F2D, (1 << 64) - 1

    def __init__(self, initstate: int = 0x853C49E6748FEA9B, initseq: int = 0xDA3E39CB94B95BDB):
        self.inc = ((initseq << 1) | 1) & self._M64
        self.state = 0
        self.next_uint32()
        self.state = (self.state + initstate) & self._M64
        self.next_uint32()

    def next_uint32(self) -> int:
        old = self.state
        self.state = (old * self._MULT + self.inc) & self._M64
        xs = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xs >> rot) | (xs << ((-rot) & 31))) & 0xFFFFFFFF

    def next_float32(self) -> float:
        return float(np.array((self.next_uint32() >> 9) | 0x3F800000, dtype=np.uint32).view(np.float32) - np.float32(1.0))


def _cuda_device(device: Optional[int]) -> torch.device:
    return torch.device("cuda", torch.cuda.current_device() if device is None else device)


def _as_grid(value, dev) -> torch.Tensor:
    """A parameter grid given as data or as the path of a `.vol` file (the reference names its volumes by
    file: `medium_filename`, `albedo_filename`, scene_config.py:102-340) -> float32 (Z, Y, X, C) on `dev`."""
    if isinstance(value, (str, os.PathLike)):
        from .multires import read_vol
        value = read_vol(os.fspath(value))[0]
    t = torch.as_tensor(value).to(device=dev, dtype=torch.float32)
    if t.dim() == 3:
        t = t.unsqueeze(-1)
    return t.contiguous().clone()


def _grid_shape(volume, key: str) -> Tuple[int, int, int, int]:
    x, y, z = volume.res
    return (z, y, x, 1 if key.endswith(SIGMA_T_SUFFIX) else 3)


def _bound_scene(volume, sigma_t_shape, device, factor: Optional[int] = None) -> Scene:
    z, y, x = sigma_t_shape[:3]
    return Scene(volume.with_resolution((x, y, z), volume.majorant_resolution_factor if factor is None else factor), device)


def reference_pass_plan(film_size: Tuple[int, int], ref_spp: int, max_rays_per_pass: int) -> Tuple[int, int]:
    """optimize.py:36-41: (pass_count, spp_per_pass) so that one pass traces at most max_rays_per_pass rays."""
    total_rays = int(film_size[0]) * int(film_size[1]) * int(ref_spp)
    pass_count = int(math.ceil(total_rays / max_rays_per_pass))
    spp_per_pass = int(math.ceil(ref_spp / pass_count))
    assert spp_per_pass * pass_count >= ref_spp
    return pass_count, spp_per_pass


def render_reference_image(scene_config, to_render: Dict[int, str], seed: int = 1234,
                           max_rays_per_pass: int = 720 * 720 * 2048, device: Optional[int] = None) -> None:
    """optimize.py:24-53: the reference medium rendered at ref_spp with `ref_integrator`, split into
    passes with consecutive seeds and averaged, one EXR per sensor."""
    if not scene_config.ref_params:
        raise ValueError(f"scene config '{scene_config.name}' has no ref_params to render references from")
    dev = _cuda_device(device)
    params = {k: _as_grid(v, dev) for k, v in scene_config.ref_params.items()}
    k_sig = next(k for k in params if k.endswith(SIGMA_T_SUFFIX))
    # the reference loads its scene file with the configuration's supergrid factor (scene_config.py:36, scene vars)
    scene = _bound_scene(scene_config.ref_volume, params[k_sig].shape, dev.index, scene_config.majorant_resolution_factor)
    integrator = load_dict({"type": scene_config.ref_integrator, "max_depth": scene_config.max_depth})
    for s, fname in to_render.items():
        sensor = scene_config.scene_sensors[s]
        pass_count, spp_per_pass = reference_pass_plan((sensor.width, sensor.height), scene_config.ref_spp, max_rays_per_pass)
        result = None
        for pass_i in range(pass_count):
            image = integrator.render(scene, params, sensor=sensor, spp=spp_per_pass, seed=seed + pass_i)
            result = image / pass_count if result is None else result + image / pass_count
        write_exr(fname, result)


def get_reference_image_paths(scene_config, overwrite: bool = False, device: Optional[int] = None) -> Dict[int, str]:
    """optimize.py:56-72: `<references>/ref_%06d.exr` per optimisation sensor, rendered when missing."""
    ref_dir = scene_config.references
    os.makedirs(ref_dir, exist_ok=True)
    paths = {s: os.path.join(ref_dir, "ref_{:06d}.exr".format(s)) for s in scene_config.sensors}
    missing = dict(paths) if overwrite else {s: f for s, f in paths.items() if not os.path.isfile(f)}
    if missing:
        render_reference_image(scene_config, missing, device=device)
    return paths


def load_reference_images(paths: Dict[int, str], batchify: bool = False, device=None):
    """optimize.py:75-88: one [n, H, W, C] tensor in the order of `paths` (ray batches) or {sensor: [H, W, C]}."""
    def rgb(f):  # the path renders RGB; an alpha channel in a foreign reference file is not compared
        return np.ascontiguousarray(read_exr(f)[..., :3])
    if batchify:
        return torch.from_numpy(np.stack([rgb(f) for f in paths.values()])).to(device)
    return {s: torch.from_numpy(rgb(f)).to(device) for s, f in paths.items()}


def initial_resolution(shape: Sequence[int], upsample) -> Tuple[int, ...]:
    """optimize.py:146-156: the resolution that reaches `shape` after len(upsample) doublings."""
    if not upsample:
        return tuple(int(v) for v in shape)
    assert len(shape) == 4
    f = 2 ** len(upsample)
    init_res = (*[max(1, int(v) // f) for v in shape[:3]], int(shape[-1]))
    if 1 in init_res[:3]:
        raise ValueError(f"Initial resolution not supported: {init_res}. Maybe reduce upsample_steps?")
    return init_res


def initialize_scene(opt_config, scene_config, device: Optional[int] = None):
    """optimize.py:134-166 -> (scene bound to the device at the initial resolution, params)."""
    from .multires import adjust_majorant_res_factor
    dev = _cuda_device(device)
    params, factor = {}, scene_config.majorant_resolution_factor
    for k, v in scene_config.start_from_value.items():
        assert k in scene_config.param_keys
        if v is None:                                            # keep what the scene file holds
            assert not opt_config.upsample
            params[k] = _as_grid(scene_config.initial_params[k], dev)
            continue
        init_res = initial_resolution(_grid_shape(scene_config.volume, k), opt_config.upsample)
        if opt_config.upsample and ".sigma_t." in k:
            factor = adjust_majorant_res_factor(scene_config.majorant_resolution_factor, init_res)
        params[k] = torch.full(init_res, float(v), dtype=torch.float32, device=dev)
    k_sig = next(k for k in params if k.endswith(SIGMA_T_SUFFIX))
    z, y, x = params[k_sig].shape[:3]
    scene = Scene(scene_config.volume.with_resolution((x, y, z), factor), dev.index)
    return scene, params


def checkpoint_prefix(opt_config, name_or_it) -> Optional[str]:
    """optimize.py:255-268: the file prefix of a checkpoint, or None when this one is not written."""
    if name_or_it == "initial":
        return "initial" if opt_config.checkpoint_initial else None
    if name_or_it == "final":
        return "final" if opt_config.checkpoint_final else None
    if isinstance(name_or_it, int) and not isinstance(name_or_it, bool):
        stride = opt_config.checkpoint_stride
        if name_or_it == 0 or not stride or name_or_it % stride != 0:
            return None
        return f"{name_or_it:08d}"
    raise ValueError("Unsupported: " + str(name_or_it))


def create_checkpoint(output_dir: str, opt_config, scene_config, params, name_or_it):
    """optimize.py:255-272: `.vol` grids of every optimised key under <output_dir>/params."""
    from .multires import save_params
    prefix = checkpoint_prefix(opt_config, name_or_it)
    if prefix is None:
        return None
    return save_params(os.path.join(output_dir, "params"), params, prefix, scene_config.param_keys)


def preview_suffix(opt_config, it_i) -> Optional[str]:
    """optimize.py:110-123."""
    if it_i == "initial":
        return "_init" if opt_config.render_initial else None
    if it_i == "final":
        return "_final" if opt_config.render_final else None
    if isinstance(it_i, int):
        return f"_{it_i:08d}"
    assert isinstance(it_i, str)
    return it_i


def render_previews(output_dir: str, opt_config, scene_config, scene: Scene, params, integrator, it_i):
    """optimize.py:110-131: `opt<suffix>_%04d.exr` of every preview sensor at seed 1234."""
    suffix = preview_suffix(opt_config, it_i)
    if suffix is None:
        return []
    preview_spp = opt_config.preview_spp or opt_config.spp
    written = []
    for s in scene_config.preview_sensors:
        fname = os.path.join(output_dir, f"opt{suffix}_{s:04d}.exr")
        image = integrator.render(scene, params, sensor=scene_config.scene_sensors[s], seed=1234, spp=preview_spp)
        write_exr(fname, image)
        written.append(fname)
    return written


def run_optimization(output_dir: str, opt_config, scene_config, int_config, device: Optional[int] = None,
                     callback: Optional[Callable[[int, float], None]] = None):
    """python/optimize.py:275-365.  Same sequence: reference images (rendered once, cached as EXR) ->
    integrator from the registry -> constant initial grids (coarse when upsampling) -> per iteration
    {seeds from the base seed, learning rates, upsampling, one ray batch or one random sensor, loss,
    backward, optimiser step + projection, medium rebuild, checkpoint, preview} -> final checkpoint.
    Returns (scene, params, opt) like the reference.  `callback(it, loss)` is an addition: the
    reference never reports its loss."""
    from .batched import gather_ref_values, render_batch
    from .multires import upsample_params
    from .opt_config import get_int_config
    os.makedirs(output_dir, exist_ok=True)
    batch_size = opt_config.batch_size
    ref_paths = get_reference_image_paths(scene_config, device=device)
    dev = _cuda_device(device)
    ref_images = load_reference_images(ref_paths, batchify=(batch_size is not None), device=dev)
    integrator = get_int_config(int_config).create(max_depth=scene_config.max_depth)
    sampler = PCG32(initstate=93483)

    n_sensors = len(scene_config.sensors)
    spp_grad = opt_config.spp
    spp_primal = spp_grad * opt_config.primal_spp_factor
    if batch_size is not None:
        batch_sensors = [scene_config.scene_sensors[s] for s in scene_config.sensors]

    scene, params = initialize_scene(opt_config, scene_config, dev.index)
    opt = opt_config.optimizer(params)

    create_checkpoint(output_dir, opt_config, scene_config, opt.params, "initial")
    render_previews(output_dir, opt_config, scene_config, scene, opt.params, integrator, "initial")
    for s in scene_config.preview_sensors:                       # the matching reference views, for comparison
        ref = ref_images[scene_config.sensors.index(s)] if batch_size is not None else ref_images[s]
        write_exr(os.path.join(output_dir, f"ref_{s:04d}.exr"), ref)

    for it_i in range(opt_config.n_iter):
        seed = _native.tea32(2 * it_i + 0, opt_config.base_seed)
        seed_grad = _native.tea32(2 * it_i + 1, opt_config.base_seed)
        opt.set_learning_rate(opt_config.learning_rates(scene_config, it_i))
        if opt_config.should_upsample(it_i):
            upsample_params(scene, opt, scene_config.majorant_resolution_factor)

        p = {k: v.requires_grad_(True) for k, v in opt.params.items()}
        if batch_size is not None:
            image, sensor_idx, pixel_idx = render_batch(batch_size, scene, batch_sensors, p, integrator, seed=seed,
                                                        seed_grad=seed_grad, spp=spp_primal, spp_grad=spp_grad)
            ref_values = gather_ref_values(ref_images, sensor_idx, pixel_idx)
        else:
            sensor_i = scene_config.sensors[int(sampler.next_float32() * n_sensors)]
            image = render(scene, p, integrator, sensor=scene_config.scene_sensors[sensor_i], spp=spp_primal,
                           spp_grad=spp_grad, seed=seed, seed_grad=seed_grad)
            ref_values = ref_images[sensor_i]
        loss_value = opt_config.loss(image, ref_values)
        loss_value.backward()
        grads = {k: v.grad for k, v in p.items()}
        for v in p.values():
            v.requires_grad_(False)
            v.grad = None

        opt.step(scene.ctx, grads, max_density=scene_config.max_density)       # opt.step() + enforce_valid_params
        scene.update_medium(opt.params[next(k for k in opt.params if k.endswith(SIGMA_T_SUFFIX))], force=True)
        create_checkpoint(output_dir, opt_config, scene_config, opt.params, it_i)
        if callback is not None:
            callback(it_i, float(loss_value.detach()))
        if it_i > 0 and it_i % opt_config.preview_stride == 0:
            render_previews(output_dir, opt_config, scene_config, scene, opt.params, integrator, it_i)

    scene.ctx.check_watchdog()
    create_checkpoint(output_dir, opt_config, scene_config, opt.params, "final")
    render_previews(output_dir, opt_config, scene_config, scene, opt.params, integrator, "final")
    return scene, opt.params, opt
