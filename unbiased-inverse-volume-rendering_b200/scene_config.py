"""Scene configurations of an optimisation run -- host mirror of python/scene_config.py:9-99.

The reference's `SceneConfig` names a Mitsuba XML file plus the variables it is loaded with
(`fname`, `normal_scene_vars`, `ref_scene_vars`); scene/XML loading is outside the hot path
(DESIGN.md §8), so here the scene itself is given as data: `volume` (a `VolumeScene`: medium box,
scale, emitter) and `scene_sensors` (what `scene.sensors()` would return), and `ref_params` holds
the grids of the reference medium (what `ref_scene_vars['medium_filename']` etc. point at).
Everything else keeps the reference's names, defaults and rules:

  * every optimised key needs an entry in `start_from_value`             (scene_config.py:54-56)
  * `references` defaults to `<OUTPUT_DIR>/references/<name>`            (scene_config.py:58-61)
  * `preview_sensors` defaults to the first optimisation sensor          (scene_config.py:63-64)
  * `param_lr_factors` defaults to 2.0 for every `.albedo.` key          (scene_config.py:67-71)
  * registry: add_scene_config / add_scene_config_variant / get_scene_config (scene_config.py:74-95)
"""
from __future__ import annotations

import copy
import dataclasses
import os
from typing import Any, Dict, List, Optional

from .scene import Sensor, VolumeScene

# python/constants.py:4-5 (the reference writes next to its sources; here: the working directory or $UIVR_OUTPUT_DIR)
OUTPUT_DIR = os.path.realpath(os.environ.get("UIVR_OUTPUT_DIR", "outputs"))


@dataclasses.dataclass
class SceneConfig:
    name: str
    volume: VolumeScene                    # stands in for fname + normal_scene_vars
    scene_sensors: List[Sensor]            # scene.sensors() of the loaded file
    param_keys: List[str]
    sensors: List[int]                     # indices into scene_sensors used by the optimisation
    start_from_value: Dict[str, Optional[float]]

    max_depth: int = 64
    references: Optional[str] = None
    ref_spp: int = 8192
    ref_integrator: str = "volpathsimple"
    ref_volume: Optional[VolumeScene] = None      # stands in for ref_fname (defaults to `volume`)
    ref_params: Optional[Dict[str, Any]] = None   # stands in for ref_scene_vars: grids of the reference medium
    initial_params: Optional[Dict[str, Any]] = None  # values of keys whose start_from_value is None (the file's own data)
    preview_sensors: Optional[List[int]] = None

    max_density: float = 250
    majorant_resolution_factor: int = 8
    param_lr_factors: Optional[Dict[str, float]] = None

    def __post_init__(self):
        for k in self.param_keys:
            if k not in self.start_from_value:
                raise ValueError(f'Parameter "{k}" will be optimized but was not given an initial value in `start_from_value`')
        for s in self.sensors:
            if not 0 <= s < len(self.scene_sensors):
                raise ValueError(f"Sensor index {s} is not part of the scene ({len(self.scene_sensors)} sensors)")
        if self.ref_volume is None:
            self.ref_volume = copy.deepcopy(self.volume)
        if self.references is None:
            self.references = os.path.join(OUTPUT_DIR, "references", self.name)
        elif not os.path.isdir(self.references):
            self.references = os.path.join(OUTPUT_DIR, "references", self.references)
        if not self.preview_sensors:
            self.preview_sensors = [self.sensors[0]]
        if not self.param_lr_factors:
            self.param_lr_factors = {k: 2.0 for k in self.param_keys if ".albedo." in k}


_SCENE_CONFIGS: Dict[str, SceneConfig] = {}
_SCENE_CONFIG_KWARGS: Dict[str, Dict[str, Any]] = {}


def add_scene_config(name: str, **kwargs) -> None:
    assert name not in _SCENE_CONFIGS, f"Duplicate scene config name: {name}"
    _SCENE_CONFIG_KWARGS[name] = dict(kwargs)
    _SCENE_CONFIGS[name] = SceneConfig(name, **kwargs)


def add_scene_config_variant(name: str, base: str, **kwargs) -> None:
    """A configuration that differs from `base` in the given fields only (scene_config.py:81-86)."""
    add_scene_config(name, **{**_SCENE_CONFIG_KWARGS[base], **kwargs})


def get_scene_config(name):
    """Always a private copy (scene_config.py:89-92).  Tensors inside are shared, not cloned."""
    out = copy.copy(name if isinstance(name, SceneConfig) else _SCENE_CONFIGS[name])
    for f in dataclasses.fields(out):
        v = getattr(out, f.name)
        if isinstance(v, (list, dict)):
            setattr(out, f.name, copy.copy(v))
    return out
