"""Benchmark workloads for the CPU arm (TEST / MEASUREMENT INFRASTRUCTURE -- not the product).

`bench.py --impl reference` and `bench.py`'s `cpu_baseline` leg time the CPU oracle on the workload the
GPU arm measures.  They must not import the product package (nor map its libuivr.so), so the workload
recipe of SURVEY §8(d) is restated here, self-contained: numpy for the scene description, torch (CPU) only
for the trilinear up-sampling of the seeded noise, the oracle's own TEA for the per-step seeds.
tests/test_host.py::test_oracle_workload_matches_product_workload keeps the two statements equal.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np

BASE_SEED = 1234

WORKLOADS = {
    # name: (grid n, film w, h, spp, dense medium, description)
    "config3": (256, 512, 512, 64, False, "BASELINE.json configs[2]"),
    "dense": (256, 512, 512, 64, True, "config3 shapes, medium without empty space (HBM-heavier case)"),
    "config5": (512, 1024, 1024, 128, False, "BASELINE.json configs[4]"),
}


def synthetic_grids(n: int, seed: int = 20220721, dense: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """SURVEY §8(d) heterogeneous recipe -> sigma_t (n,n,n,1) in [0,1], albedo (n,n,n,3), float32.
    dense: no spherical fall-off and no `< 0.05 -> 0` cut: d = 0.5 + 0.5 f, so that every supergrid cell is
    occupied and the optical thickness across the box is ~12 at scale 8."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(1, 1, 16, 16, 16, generator=g)
    f = F.interpolate(c, size=(n, n, n), mode="trilinear", align_corners=True)[0, 0]
    if dense:
        d = 0.5 + 0.5 * f
    else:
        ax = (torch.arange(n, dtype=torch.float32) + 0.5) / n - 0.5
        r2 = ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2
        fall = torch.clamp(1.0 - r2 / 0.25, 0.0, 1.0)
        d = f * f * fall
        d = d / d.max()
        d[d < 0.05] = 0.0
    a = torch.rand(1, 3, 8, 8, 8, generator=g)
    a = F.interpolate(a, size=(n, n, n), mode="trilinear", align_corners=True)[0]
    albedo = (0.2 + 0.75 * a).permute(1, 2, 3, 0).contiguous()
    return d.unsqueeze(-1).contiguous().numpy(), albedo.numpy()


def benchmark_desc(n: int, width: int, height: int, scale: float = 8.0, factor: int = 8) -> Dict[str, object]:
    """Scene description (the oracle's `desc`) of SURVEY §8(d): medium box [-0.5, 1.5]^3, perspective sensor
    fov_x 30 deg at (4,4,4) looking at the box centre, constant emitter (1, 0.8, 0.2), supergrid factor 8."""
    origin = np.array([4.0, 4.0, 4.0])
    target = np.array([0.5, 0.5, 0.5])
    up = np.array([0.0, 1.0, 0.0])
    d = target - origin
    d /= np.linalg.norm(d)
    left = np.cross(up, d)
    left /= np.linalg.norm(left)
    new_up = np.cross(d, left)
    tan_x = math.tan(math.radians(30.0) * 0.5)
    f = int(factor)
    while f > 1 and (n // f) < 4:   # optimize.py:182-199 adjust_majorant_res_factor
        f -= 1
    to_local = np.zeros((3, 4))
    for a in range(3):
        to_local[a, a] = 0.5
        to_local[a, 3] = 0.25
    return {
        "res": (n, n, n), "to_local": to_local.astype(np.float32).reshape(-1), "scale": np.float32(scale),
        "majorant_factor": 0 if f <= 1 else f, "radiance": np.array([1.0, 0.8, 0.2], dtype=np.float32),
        "local_to_world": np.diag([2.0, 2.0, 2.0]).astype(np.float32).reshape(-1),
        "cam_origin": origin.astype(np.float32), "cam_left": left.astype(np.float32),
        "cam_up": new_up.astype(np.float32), "cam_dir": d.astype(np.float32),
        "tan_x": np.float32(tan_x), "tan_y": np.float32(tan_x * height / width), "near_clip": np.float32(1e-2),
        "width": int(width), "height": int(height),
    }


def drt_props(max_depth: int = 64) -> Dict[str, object]:
    """`volpathsimple-drt` of opt_config.py:133-142 as created by IntegratorConfig.create(max_depth=64)."""
    return dict(max_depth=max_depth, hide_emitters=False, use_nee=True, use_drt=True, use_drt_subsampling=True,
                use_drt_mis=True)


def step_seeds(O, it: int) -> Tuple[int, int]:
    """optimize.py:327-328: seed, seed_grad = tea32(2 it, base), tea32(2 it + 1, base)."""
    return O.tea(2 * it, BASE_SEED)[0], O.tea(2 * it + 1, BASE_SEED)[0]


def oracle_step(O, desc, props, sig, alb, it: int, spp: int, nthreads: int):
    """One step of the metric on the CPU: forward at `seed`, loss gradient, DRT backward at `seed_grad`."""
    seed, seed_grad = step_seeds(O, it)
    img, _, _ = O.render_forward(desc, props, sig, alb, seed, spp, nthreads=nthreads)
    g = (2.0 * (img.astype(np.float64) - 0.5) / img.size).astype(np.float32)
    O.render_backward(desc, props, sig, alb, g, seed_grad, spp, nthreads=nthreads)
