"""Host mirror of the reference's `IntegratorConfig` surface (python/opt_config.py:83-169).

Same field names, registry names, property keys and error behaviour (AssertionError on the
same preconditions), so that `get_int_config(name).create(max_depth=...)` -- the call made at
python/optimize.py:290 -- returns the native-backed integrator where the reference returns
`mi.load_dict(...)`.  Written table-first: the registry below is data, `create` is the only
logic.
"""
from __future__ import annotations

import copy
import dataclasses
import enum
from typing import Any, Callable, Dict, List, Optional, Union

from . import losses
from .integrator import load_dict


class Schedule(enum.IntEnum):
    """opt_config.py:78-80."""
    Constant = 0
    Last25 = 1


@dataclasses.dataclass
class OptimizationConfig:
    """opt_config.py:11-75: the options of one optimisation run, same field names and defaults.
    `optimizer(params)` hands out the native-backed optimiser (optimize.Adam: fused update +
    projection kernel) where the reference constructs `mi.ad.Adam` / `mi.ad.SGD`."""
    # required, in the reference's positional order (opt_config.py:14-17)
    name: str
    spp: int                                       # samples per pixel of the adjoint pass
    n_iter: int
    lr: float
    # defaults, in the reference's positional order as well (opt_config.py:19-37): reference-style positional
    # construction binds the same fields
    primal_spp_factor: int = 64                    # primal spp = spp x this
    batch_size: Optional[int] = None               # None: one random sensor per iteration; else rays per batch
    lr_schedule: Optional[Schedule] = None
    upsample: Optional[List[float]] = None         # fractions of the run at which the grids double
    base_seed: int = 988378
    render_initial: bool = True
    render_final: bool = True
    preview_stride: int = 100
    checkpoint_initial: bool = True
    checkpoint_final: bool = True
    checkpoint_stride: Optional[int] = 1000        # None / 0: no intermediate checkpoints
    preview_spp: Optional[int] = None              # None: spp
    opt_type: str = "adam"                         # "adam" | "sgd"
    opt_args: Optional[Dict[str, Any]] = None
    loss: Callable = losses.l1

    def __post_init__(self):
        from .multires import upsample_iterations
        self.upsample_at = upsample_iterations(self.upsample, self.n_iter)          # opt_config.py:39-44

    def optimizer(self, params):
        """opt_config.py:46-48; an unknown opt_type is a KeyError there as well."""
        from . import optimize
        make = {"sgd": optimize.SGD, "adam": optimize.Adam}[self.opt_type]
        return make(lr=self.lr, params=params, **(self.opt_args or {}))

    def learning_rates(self, scene_config, it_i: int) -> Dict[str, float]:
        """opt_config.py:50-69, through optimize.learning_rates (which holds the Last25 rule)."""
        from .optimize import learning_rates
        if self.lr_schedule is not None and self.lr_schedule not in tuple(Schedule):
            raise ValueError(f"Unsupported schedule: {self.lr_schedule}")
        name = None if self.lr_schedule is None else Schedule(self.lr_schedule).name.lower()
        return learning_rates(self.lr, scene_config.param_keys, it_i, self.n_iter, name, scene_config.param_lr_factors)

    def should_upsample(self, it_i: int) -> bool:
        """opt_config.py:72-75."""
        return it_i in self.upsample_at


# rr_depth = max_depth + this: Russian roulette never fires (opt_config.py:103-106)
_RR_OFFSET = 1000


@dataclasses.dataclass
class IntegratorConfig:
    """opt_config.py:83-95.  `params` is the plugin dictionary handed to `load_dict`."""
    name: str
    pretty_name: str
    params: Dict[str, Any]
    uses_fd: bool = False
    fd_epsilon: Optional[float] = None
    fd_spp_multiplier: int = 16

    def __post_init__(self):
        assert not self.uses_fd or self.fd_epsilon is not None, "finite differences need fd_epsilon"

    def create(self, **overrides):
        """opt_config.py:97-108: merge overrides, require max_depth, forbid a caller rr_depth."""
        assert "max_depth" in overrides, "create() requires max_depth"
        assert "rr_depth" not in overrides, "rr_depth is not configurable (Russian roulette is unsupported)"
        plugin = {**copy.deepcopy(self.params), **overrides}
        assert plugin["max_depth"] >= 0
        plugin.setdefault("rr_depth", plugin["max_depth"] + _RR_OFFSET)
        return load_dict(plugin)


_REGISTRY: Dict[str, IntegratorConfig] = {}


def add_int_config(name: str, **fields) -> None:
    """opt_config.py:112-114."""
    assert name not in _REGISTRY, f"Duplicate integrator config name: {name}"
    _REGISTRY[name] = IntegratorConfig(name, **fields)


def get_int_config(name: Union[str, IntegratorConfig]) -> IntegratorConfig:
    """opt_config.py:117-120: always hands out a private copy."""
    src = name if isinstance(name, IntegratorConfig) else _REGISTRY[name]
    return copy.deepcopy(src)


def _vps(**flags) -> Dict[str, Any]:
    return {"type": "volpathsimple", **flags}


# opt_config.py:123-169 -- (registry name, pretty name, plugin params, extra fields)
for _name, _pretty, _params, _extra in (
    ("fd-forward", "Finite differences", _vps(use_drt=False), {"uses_fd": True, "fd_epsilon": 5e-3}),
    ("volpathsimple-drt", "Differential Ratio Tracking",
     _vps(use_drt=True, use_drt_subsampling=True, use_drt_mis=True), {}),
    ("volpathsimple-drt-quadratic", "Differential Ratio Tracking (quadratic)",
     _vps(use_drt=True, use_drt_subsampling=False, use_drt_mis=True), {}),
    ("volpathsimple-basic", "Free-flight based", _vps(use_drt=False), {}),
    # emission-absorption ray marcher (python/integrators/nerf.py), the reference's warm-start integrator
    ("nerf", "NeRF (grid-backed)", {"type": "nerf", "queries_per_ray": 128}, {}),
):
    add_int_config(_name, pretty_name=_pretty, params=_params, **_extra)
