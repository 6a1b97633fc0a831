"""Host-side mirror of the reference's integrator plugin surface for the volpathsimple path.

Reference interface                                  -> here
  mi.register_integrator("volpathsimple", ...)  (volpathsimple.py:769)   -> INTEGRATORS registry
  VolpathSimpleIntegrator(props)                 (volpathsimple.py:19-34) -> VolpathSimpleIntegrator
  RBIntegrator.render / render_backward          (batched.py:134-326)     -> .render / .render_backward
  mi.render(scene, params, integrator, sensor, spp, spp_grad, seed, seed_grad)
                                                 (optimize.py:345-347)    -> render(...)
  dr.backward(loss) -> dr.grad(params[k])        (optimize.py:350)        -> torch autograd

PyTorch only holds device memory and streams; all arithmetic runs in libuivr.so (C-ABI,
include/uivr.h).  There is no CPU or eager fallback.
"""
from __future__ import annotations

import weakref

from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _native
from .scene import Sensor, VolumeScene

SIGMA_T_SUFFIX = "sigma_t.data"
ALBEDO_SUFFIX = "albedo.data"
EMISSION_SUFFIX = "emission.data"


def _find_key(params: Dict[str, torch.Tensor], suffix: str) -> str:
    keys = [k for k in params.keys() if k.endswith(suffix)]
    if len(keys) != 1:
        # util.get_single_medium (util.py:75-86): exactly one medium / one grid of each kind
        raise ValueError(f"expected exactly one parameter ending in '{suffix}', found {keys}")
    return keys[0]


def _stream() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


class Scene:
    """Runtime scene: a VolumeScene bound to one CUDA device through a native context."""

    def __init__(self, volume: VolumeScene, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise _native.NativeError("CUDA device required: the volpathsimple path has no CPU fallback")
        self.volume = volume
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.ctx = _native.Context(self.device)
        self._medium_key = None
        self._scene_key = None
        self._props_key = None
        self._env_ref = None

    # mi.traverse(scene) equivalent for the two optimised grids
    def check_params(self, params: Dict[str, torch.Tensor],
                     second_suffix: str = ALBEDO_SUFFIX) -> Tuple[torch.Tensor, torch.Tensor]:
        x, y, z = self.volume.res
        sig = params[_find_key(params, SIGMA_T_SUFFIX)]
        alb = params[_find_key(params, second_suffix)]
        for name, t, shape in (("sigma_t", sig, (z, y, x, 1)), (second_suffix.split(".")[0], alb, (z, y, x, 3))):
            if tuple(t.shape) != shape and not (name == "sigma_t" and tuple(t.shape) == (z, y, x)):
                raise ValueError(f"{name} must have shape {shape}, got {tuple(t.shape)}")
            if t.dtype != torch.float32 or not t.is_cuda or t.device.index != self.device:
                raise ValueError(f"{name} must be a float32 CUDA tensor on device {self.device}")
            if not t.is_contiguous():
                raise ValueError(f"{name} must be contiguous")
        return sig, alb

    def bind(self, sensor: Optional[Sensor], props: dict):
        vol = self.volume if sensor is None else self.volume.with_sensor(sensor)
        desc = vol.as_dict()
        env_keys = ("env_data", "env_marg", "env_cond")  # big tables: keyed by identity of the EnvMap object
        key = tuple((k, tuple(map(float, v.reshape(-1))) if hasattr(v, "reshape") else v)
                    for k, v in sorted(desc.items()) if k not in env_keys)
        if key != self._scene_key:
            old_medium_inputs = None if self._scene_key is None else self._medium_inputs
            self.ctx.set_scene(desc)
            # The environment map has its own key: a sensor change (one per iteration in the optimisation loop) must
            # not re-upload tens of MB of tables.  The strong reference keeps id() from being recycled.
            if self._env_ref is not vol.envmap or self._scene_key is None:
                self.ctx.set_envmap(desc)
                self._env_ref = vol.envmap
            self._scene_key = key
            self._medium_inputs = (desc["res"], float(desc["scale"]), desc["majorant_factor"])
            if old_medium_inputs != self._medium_inputs:
                self._medium_key = None
        pkey = None if props is None else tuple(sorted(props.items()))
        if pkey is not None and pkey != self._props_key:
            self.ctx.set_integrator(props)
            self._props_key = pkey
        return desc

    def update_medium(self, sigma_t: torch.Tensor, force: bool = False):
        """params.update(): rebuild the octet copy of sigma_t, the majorant supergrid and the walk table unless
        this very tensor OBJECT, unmodified, is what they were built from.  The key must not alias: a fresh
        tensor (params[k] = softplus(raw), fd-style params[k] = new) has version 0 and usually gets the freed
        address of its predecessor back from the caching allocator, so (data_ptr, _version) is not enough --
        identity is tracked with a weak reference.  Code that writes through raw pointers (optimize.Adam ->
        uivr_adam_step) bumps the version counter itself (torch.autograd.graph.increment_version)."""
        ref = self._medium_key[0]() if self._medium_key is not None else None
        if force or ref is not sigma_t or self._medium_key[1:] != (sigma_t.data_ptr(), sigma_t._version):
            self.ctx.update_medium(sigma_t.data_ptr(), _stream())
            self._medium_key = (weakref.ref(sigma_t), sigma_t.data_ptr(), sigma_t._version)


    def update_medium_after_reshape(self, sigma_t: torch.Tensor):
        """The medium changed resolution (self.volume was replaced): re-describe the scene to the
        native context and rebuild its lookup structures on the next render."""
        self._scene_key = None
        self._medium_key = None


class VolpathSimpleIntegrator:
    """Same property names / defaults as python/integrators/volpathsimple.py:19-34."""

    second_suffix = ALBEDO_SUFFIX  # the RGB grid differentiated next to sigma_t

    def __init__(self, props: Optional[dict] = None):
        props = dict(props or {})
        props.pop("type", None)
        self.max_depth = int(props.pop("max_depth", 64))
        self.rr_depth = int(props.pop("rr_depth", self.max_depth + 1000))
        self.hide_emitters = bool(props.pop("hide_emitters", False))
        self.use_nee = bool(props.pop("use_nee", True))
        self.use_drt = bool(props.pop("use_drt", True))
        self.use_drt_subsampling = bool(props.pop("use_drt_subsampling", True))
        self.use_drt_mis = bool(props.pop("use_drt_mis", True))
        if props:
            raise ValueError(f"unknown integrator properties: {sorted(props)}")
        if self.max_depth < 0:
            raise ValueError("max_depth must be >= 0")
        if self.rr_depth <= self.max_depth:
            # volpathsimple.py:36 "TODO: support Russian Roulette"; opt_config.py:103-106
            raise NotImplementedError("Russian roulette is not supported (rr_depth must exceed max_depth)")

    def props(self) -> dict:
        return {"max_depth": self.max_depth, "hide_emitters": self.hide_emitters,
                "use_nee": self.use_nee, "use_drt": self.use_drt,
                "use_drt_subsampling": self.use_drt_subsampling, "use_drt_mis": self.use_drt_mis}

    def aovs(self):
        return []

    # RBIntegrator.render (batched.py:134-197)
    def render(self, scene: Scene, params: Dict[str, torch.Tensor], sensor: Optional[Sensor] = None,
               seed: int = 0, spp: int = 0, develop: bool = True, evaluate: bool = True,
               shard: Optional[Sequence[int]] = None, sample_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not develop:
            raise Exception("develop=True must be specified when invoking AD integrators")  # batched.py:145-147
        if spp <= 0:
            raise ValueError("spp must be positive")
        sig, alb = scene.check_params(params)
        desc = scene.bind(sensor, self.props())
        scene.update_medium(sig.detach())
        image = torch.empty((desc["height"], desc["width"], 3), dtype=torch.float32, device=sig.device)
        scene.ctx.render_forward(alb.detach().data_ptr(), seed, spp, image.data_ptr(),
                                 None if sample_out is None else sample_out.data_ptr(), shard, _stream())
        return image

    # RBIntegrator.render_backward (batched.py:212-326)
    def render_backward(self, scene: Scene, params: Dict[str, torch.Tensor], grad_in: torch.Tensor,
                        sensor: Optional[Sensor] = None, seed: int = 0, spp: int = 0,
                        shard: Optional[Sequence[int]] = None, sample_out: Optional[torch.Tensor] = None,
                        out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        if spp <= 0:
            raise ValueError("spp must be positive")
        sig, alb = scene.check_params(params)
        desc = scene.bind(sensor, self.props())
        if tuple(grad_in.shape) != (desc["height"], desc["width"], 3):
            raise ValueError(f"grad_in must have shape {(desc['height'], desc['width'], 3)}")
        grad_in = grad_in.to(dtype=torch.float32).contiguous()
        scene.update_medium(sig.detach())
        if out is None:
            dsig = torch.empty_like(sig)
            dalb = torch.empty_like(alb)
        else:
            dsig, dalb = out
        scene.ctx.render_backward(alb.detach().data_ptr(), grad_in.data_ptr(), seed, spp,
                                  dsig.data_ptr(), dalb.data_ptr(),
                                  None if sample_out is None else sample_out.data_ptr(), shard, _stream())
        return dsig, dalb


class NeRFIntegrator:
    """python/integrators/nerf.py: emission-absorption ray marching over the sigma_t grid and an RGB
    emission grid (`...emission.data`, (Z,Y,X,3)); same property names / defaults as nerf.py:27-35."""

    second_suffix = EMISSION_SUFFIX

    def __init__(self, props: Optional[dict] = None):
        props = dict(props or {})
        props.pop("type", None)
        # RBIntegrator base-class properties (unused by this integrator, accepted like the reference)
        self.max_depth = int(props.pop("max_depth", 6))
        self.rr_depth = int(props.pop("rr_depth", 5))
        self.hide_emitters = bool(props.pop("hide_emitters", False))
        self.queries_per_ray = int(props.pop("queries_per_ray", 128))
        self.density_noise_std = float(props.pop("density_noise_std", 0.0))
        self.jittering_enabled = bool(props.pop("jittering_enabled", True))
        self.activation_type = str(props.pop("activation", "identity")).lower()
        if props:
            raise ValueError(f"unknown integrator properties: {sorted(props)}")
        if self.density_noise_std > 0:
            # nerf.py:156-158: "Incorrect for now: noise rnd is wrong on second loop of adjoint"
            raise NotImplementedError("density_noise_std > 0 is marked incorrect by the reference (nerf.py:157)")

    def props(self) -> dict:
        if self.activation_type not in ("identity", "relu"):
            raise ValueError(f"Unsupported activation: {self.activation_type}")  # nerf.py:44
        return {"queries_per_ray": self.queries_per_ray, "jittering_enabled": self.jittering_enabled,
                "activation": self.activation_type, "hide_emitters": self.hide_emitters}

    def aovs(self):
        return []

    def render(self, scene: Scene, params: Dict[str, torch.Tensor], sensor: Optional[Sensor] = None,
               seed: int = 0, spp: int = 0, develop: bool = True, evaluate: bool = True,
               shard: Optional[Sequence[int]] = None, sample_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not develop:
            raise Exception("develop=True must be specified when invoking AD integrators")  # batched.py:145-147
        if spp <= 0:
            raise ValueError("spp must be positive")
        props = self.props()
        sig, em = scene.check_params(params, EMISSION_SUFFIX)
        desc = scene.bind(sensor, None)
        scene.update_medium(sig.detach())
        image = torch.empty((desc["height"], desc["width"], 3), dtype=torch.float32, device=sig.device)
        scene.ctx.nerf_forward(props, em.detach().data_ptr(), seed, spp, image.data_ptr(),
                               None if sample_out is None else sample_out.data_ptr(), shard, _stream())
        return image

    def render_backward(self, scene: Scene, params: Dict[str, torch.Tensor], grad_in: torch.Tensor,
                        sensor: Optional[Sensor] = None, seed: int = 0, spp: int = 0,
                        shard: Optional[Sequence[int]] = None, sample_out: Optional[torch.Tensor] = None,
                        out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        if spp <= 0:
            raise ValueError("spp must be positive")
        props = self.props()
        sig, em = scene.check_params(params, EMISSION_SUFFIX)
        desc = scene.bind(sensor, None)
        if tuple(grad_in.shape) != (desc["height"], desc["width"], 3):
            raise ValueError(f"grad_in must have shape {(desc['height'], desc['width'], 3)}")
        grad_in = grad_in.to(dtype=torch.float32).contiguous()
        scene.update_medium(sig.detach())
        dsig, dem = (torch.empty_like(sig), torch.empty_like(em)) if out is None else out
        scene.ctx.nerf_backward(props, em.detach().data_ptr(), grad_in.data_ptr(), seed, spp, dsig.data_ptr(),
                                dem.data_ptr(), None if sample_out is None else sample_out.data_ptr(), shard,
                                _stream())
        return dsig, dem


INTEGRATORS = {"volpathsimple": VolpathSimpleIntegrator, "nerf": NeRFIntegrator}


def register_integrator(name: str, factory):
    """mi.register_integrator (volpathsimple.py:769)."""
    INTEGRATORS[name] = factory


def load_dict(d: dict):
    """mi.load_dict for integrator dictionaries (opt_config.py:108)."""
    if "type" not in d:
        raise ValueError("integrator dictionary needs a 'type'")
    if d["type"] not in INTEGRATORS:
        raise NotImplementedError(f"integrator type '{d['type']}' is not part of this build")
    return INTEGRATORS[d["type"]](d)


class _RenderOp(torch.autograd.Function):
    """mi.render's _RenderOp: forward = primal render at `seed`/`spp` (detached), backward =
    render_backward at `seed_grad`/`spp_grad`."""

    @staticmethod
    def forward(ctx, sigma_t, albedo, scene, integrator, sensor, seed, seed_grad, spp, spp_grad, keys, shard, reducer):
        params = {keys[0]: sigma_t, keys[1]: albedo}
        image = integrator.render(scene, params, sensor=sensor, seed=seed, spp=spp, shard=shard)
        ctx.save_for_backward(sigma_t, albedo)
        ctx.meta = (scene, integrator, sensor, seed_grad, spp_grad, keys, shard, reducer)
        return image

    @staticmethod
    def backward(ctx, grad_image):
        sigma_t, albedo = ctx.saved_tensors
        scene, integrator, sensor, seed_grad, spp_grad, keys, shard, reducer = ctx.meta
        params = {keys[0]: sigma_t, keys[1]: albedo}
        dsig, dalb = integrator.render_backward(scene, params, grad_image, sensor=sensor, seed=seed_grad,
                                                spp=spp_grad, shard=shard)
        if reducer is not None:
            dsig, dalb = reducer(dsig, dalb)
        return (dsig, dalb) + (None,) * 10


def render(scene: Scene, params: Dict[str, torch.Tensor], integrator: VolpathSimpleIntegrator,
           sensor: Optional[Sensor] = None, spp: int = 0, spp_grad: int = 0, seed: int = 0,
           seed_grad: int = 0, shard: Optional[Sequence[int]] = None, reducer=None) -> torch.Tensor:
    """mi.render(scene, params=..., integrator=..., sensor=..., spp=..., spp_grad=..., seed=...,
    seed_grad=...) as called at optimize.py:345-347.  Differentiable w.r.t. the two grids."""
    if spp_grad == 0:
        spp_grad = spp
    if seed_grad == 0:
        # de-correlate the primal and differential phases (batched.py:119-121)
        seed_grad = _native.tea32(seed, 1)
    elif seed_grad == seed:
        raise Exception("The primal and differential seed should be different "
                        "to ensure unbiased gradient computation!")  # batched.py:122-124
    k_sig, k_alb = _find_key(params, SIGMA_T_SUFFIX), _find_key(params, integrator.second_suffix)
    return _RenderOp.apply(params[k_sig], params[k_alb], scene, integrator, sensor, seed, seed_grad,
                           spp, spp_grad, (k_sig, k_alb), shard, reducer)
