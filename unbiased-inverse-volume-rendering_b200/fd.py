"""Finite-difference gradients -- host mirror of python/fd.py:9-69 (`fd_gradients`), the harness
the reference's own gradient tests are built on (tests/test_integrators.py:170-176, :296-300) and
what the `fd-forward` IntegratorConfig (`uses_fd`, `fd_epsilon`, `fd_spp_multiplier`;
python/opt_config.py:123-132) stands for.

Forward differences with a common seed: one render at the centre, then one render per parameter
ENTRY with that entry raised by `eps` (params.update() in between, i.e. the medium's lookup
structures are rebuilt), `(loss_offset - loss_center) / eps`.  Every render goes through the
drop-in integrator surface, i.e. the CUDA path.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np
import torch

from .integrator import Scene
from .scene import Sensor


def fd_gradients(scene: Scene, params: Dict[str, torch.Tensor], loss_fn: Callable[[torch.Tensor], torch.Tensor],
                 eps: float, spp: int = 4096, integrator=None, seed: int = 1234,
                 sensor: Optional[Sensor] = None) -> Dict[str, np.ndarray]:
    """-> {key: array of d loss / d entry, same shape as the parameter} (fd.py returns the same dict).
    `output_dir` / `write_images` of the reference (EXR dumps) are not mirrored."""
    if integrator is None:
        raise ValueError("fd_gradients needs an integrator (the scene description carries none)")
    with torch.no_grad():
        loss_center = float(loss_fn(integrator.render(scene, params, sensor=sensor, seed=seed, spp=spp)))
        results = {}
        for k in list(params.keys()):
            p = params[k]
            flat = p.view(-1)
            grads = np.full(tuple(p.shape), np.nan)
            for i in range(flat.numel()):
                original = flat[i].clone()
                flat[i] = original + eps                       # fd.py:38-46 (one entry at a time)
                loss_offset = float(loss_fn(integrator.render(scene, params, sensor=sensor, seed=seed, spp=spp)))
                grads[np.unravel_index(i, grads.shape)] = (loss_offset - loss_center) / eps
                flat[i] = original                             # fd.py:65-67 restore
            results[k] = grads
    return results
