"""IntegratorConfig surface of the reference (python/opt_config.py:83-169), same names / keys /
error behaviour; `create()` returns the native-backed integrator instead of `mi.load_dict`."""
from __future__ import annotations

from copy import deepcopy
from dataclasses import dataclass
from typing import Dict, Optional

from .integrator import load_dict


@dataclass
class IntegratorConfig:
    name: str
    pretty_name: str
    params: Dict

    uses_fd: bool = False
    fd_epsilon: Optional[float] = None
    fd_spp_multiplier: int = 16

    def __post_init__(self):
        if self.uses_fd:
            assert self.fd_epsilon is not None

    def create(self, **kwargs):
        assert 'max_depth' in kwargs
        d = deepcopy(self.params)
        d.update(kwargs)

        assert d['max_depth'] >= 0
        # Russian roulette is unsupported: it never fires (opt_config.py:103-106)
        assert 'rr_depth' not in kwargs
        if 'rr_depth' not in self.params:
            d['rr_depth'] = d['max_depth'] + 1000

        return load_dict(d)


_INTEGRATOR_CONFIGS: Dict[str, IntegratorConfig] = {}


def add_int_config(name, **kwargs):
    assert name not in _INTEGRATOR_CONFIGS, f'Duplicate integrator config name: {name}'
    _INTEGRATOR_CONFIGS[name] = IntegratorConfig(name, **kwargs)


def get_int_config(name):
    if isinstance(name, IntegratorConfig):
        return deepcopy(name)
    return deepcopy(_INTEGRATOR_CONFIGS[name])


add_int_config('fd-forward', pretty_name='Finite differences',
               params={'type': 'volpathsimple', 'use_drt': False},
               uses_fd=True, fd_epsilon=5e-3)
add_int_config('volpathsimple-drt', pretty_name='Differential Ratio Tracking',
               params={'type': 'volpathsimple', 'use_drt': True, 'use_drt_subsampling': True,
                       'use_drt_mis': True})
add_int_config('volpathsimple-drt-quadratic', pretty_name='Differential Ratio Tracking (quadratic)',
               params={'type': 'volpathsimple', 'use_drt': True, 'use_drt_subsampling': False,
                       'use_drt_mis': True})
add_int_config('volpathsimple-basic', pretty_name='Free-flight based',
               params={'type': 'volpathsimple', 'use_drt': False})
# the emission-only ray marcher is a different estimator outside this build's hot path
add_int_config('nerf', pretty_name='NeRF (grid-backed)',
               params={'type': 'nerf', 'queries_per_ray': 128})
