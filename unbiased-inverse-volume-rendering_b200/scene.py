"""Scene description for the volpathsimple hot path.

The reference describes its scenes as Mitsuba dictionaries / XML (one convex medium with a
`null` boundary, one perspective sensor with a box-filter hdrfilm, one infinite emitter:
python/integrators/volpathsimple.py:11-17, tests/test_integrators.py:19-116).  Only the
quantities the path actually consumes are kept here; everything is resolved on the host in
double precision and handed to the native library as float32 (`as_dict()` is the interchange
format, also consumed by the test oracle).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Sequence, Tuple

import numpy as np


def _normalize(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v)


def look_at(origin: Sequence[float], target: Sequence[float], up: Sequence[float]):
    """Mitsuba `Transform4f.look_at` frame: columns (left, new_up, dir) and the origin."""
    origin = np.asarray(origin, dtype=np.float64)
    d = _normalize(np.asarray(target, dtype=np.float64) - origin)
    left = _normalize(np.cross(np.asarray(up, dtype=np.float64), d))
    new_up = np.cross(d, left)
    return origin, left, new_up, d


@dataclass
class Sensor:
    """Perspective sensor (fov along x) + box-filter film, tests/test_integrators.py:46-66."""
    origin: Tuple[float, float, float] = (4.0, 4.0, 4.0)
    target: Tuple[float, float, float] = (0.0, -0.15, 0.0)
    up: Tuple[float, float, float] = (0.0, 1.0, 0.0)
    fov: float = 30.0
    near_clip: float = 1e-2
    width: int = 128
    height: int = 128

    def frame(self) -> Dict[str, object]:
        o, left, up, d = look_at(self.origin, self.target, self.up)
        tan_x = math.tan(math.radians(self.fov) * 0.5)
        return {
            "cam_origin": o.astype(np.float32), "cam_left": left.astype(np.float32),
            "cam_up": up.astype(np.float32), "cam_dir": d.astype(np.float32),
            "tan_x": np.float32(tan_x),
            "tan_y": np.float32(tan_x * self.height / self.width),
            "near_clip": np.float32(self.near_clip),
            "width": int(self.width), "height": int(self.height),
        }


def circle_sensors(n: int, width: int, height: int, center=(0.5, 0.5, 0.5), radius=6.93,
                   elevation_deg=35.26, fov=30.0):
    """n cameras on a circle around the medium."""
    c = np.asarray(center, dtype=np.float64)
    el = math.radians(elevation_deg)
    out = []
    for k in range(n):
        phi = 2.0 * math.pi * k / n + math.pi / 4.0
        o = c + radius * np.array([math.cos(phi) * math.cos(el), math.sin(el),
                                   math.sin(phi) * math.cos(el)])
        out.append(Sensor(origin=tuple(o), target=tuple(c), fov=fov, width=width, height=height))
    return out


@dataclass
class EnvMap:
    """Lat-long environment emitter (Mitsuba `envmap`; every scene of python/scene_config.py:102-340
    uses one): `image` (H, W, 3) float32, row 0 = +Y pole, radiance = `scale` x bilinear lookup.

    Conventions (upstream envmap.cpp as recalled in SURVEY App. B, defined here where it cannot be
    checked): a direction d in the emitter's frame has texture coordinates
    u = atan2(d.x, -d.z) / 2pi - 0.5 / W (wrapped), v = acos(d.y) / pi; the texel grid is
    (H, W + 1) vertices (first column repeated at the end) joined by W x (H - 1) bilinear patches.
    Importance sampling: piecewise-constant density over the patches, proportional to the patch mean
    of luminance x sin(theta) (marginal over rows, conditional over columns; tables built here in
    float64 and shared verbatim by the CUDA path and the test oracle)."""
    image: np.ndarray
    scale: float = 1.0
    to_world: Tuple[Tuple[float, float, float], ...] = ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0))

    @classmethod
    def from_file(cls, path: str, scale: float = 1.0, to_world=None) -> "EnvMap":
        """Mitsuba's `envmap` `filename` (`envmap_filename` of python/scene_config.py:102-340): a Radiance
        `.hdr` or an OpenEXR lat-long image; negative / non-finite texels (seen in captured maps) become 0."""
        if path.lower().endswith((".hdr", ".rgbe", ".pic")):
            from .rgbe import read_hdr
            img = read_hdr(path)
        elif path.lower().endswith(".exr"):
            from .exr import read_exr
            img = read_exr(path)[..., :3]
        else:
            raise ValueError(f"unsupported environment map format: {path}")
        img = np.where(np.isfinite(img) & (img > 0), img, 0.0).astype(np.float32)
        return cls(img, scale) if to_world is None else cls(img, scale, to_world)

    def tables(self) -> Dict[str, object]:
        """Built once per EnvMap object (the image is treated as immutable)."""
        cached = getattr(self, "_tables", None)
        if cached is None:
            cached = self._build_tables()
            object.__setattr__(self, "_tables", cached)
        return cached

    def _build_tables(self) -> Dict[str, object]:
        img = np.asarray(self.image, dtype=np.float64)
        if img.ndim != 3 or img.shape[2] != 3 or img.shape[0] < 2 or img.shape[1] < 1:
            raise ValueError("envmap image must have shape (H >= 2, W >= 1, 3)")
        if not np.all(np.isfinite(img)) or np.any(img < 0):
            raise ValueError("envmap radiance must be finite and non-negative")
        h, w = img.shape[:2]
        verts = np.concatenate([img, img[:, :1]], axis=1)                      # (H, W+1, 3)
        lum = verts @ np.array([0.212671, 0.715160, 0.072169])                 # Rec.709 luminance
        theta = np.arange(h, dtype=np.float64) * (np.pi / (h - 1))
        f = lum * np.sin(theta)[:, None]
        patch = 0.25 * (f[:-1, :-1] + f[:-1, 1:] + f[1:, :-1] + f[1:, 1:])     # (H-1, W)
        total = patch.sum()
        if not total > 0:
            raise ValueError("envmap is black: nothing to importance-sample")
        pdf = patch * (patch.size / total)                                     # density over [0,1]^2
        rows = patch.sum(axis=1)
        marg = np.cumsum(rows) / total
        marg[-1] = 1.0
        cond = np.cumsum(patch, axis=1)
        cond = np.where(rows[:, None] > 0, cond / np.maximum(rows[:, None], 1e-300),
                        (np.arange(w) + 1.0)[None, :] / w)
        cond[:, -1] = 1.0
        data = np.zeros((h, w + 1, 4), dtype=np.float32)
        data[..., :3] = verts
        data[:-1, :-1, 3] = pdf                                                # patch (y, x) -> its density
        r = np.asarray(self.to_world, dtype=np.float64).reshape(3, 3)
        return {"env_data": np.ascontiguousarray(data), "env_w": int(w), "env_h": int(h),
                "env_scale": np.float32(self.scale),
                "env_marg": np.ascontiguousarray(marg, dtype=np.float32),
                "env_cond": np.ascontiguousarray(cond, dtype=np.float32),
                "env_to_world": r.astype(np.float32).reshape(-1),
                "world_to_env": np.linalg.inv(r).astype(np.float32).reshape(-1)}


@dataclass
class VolumeScene:
    """One heterogeneous medium in a box + one sensor + a constant emitter."""
    res: Tuple[int, int, int]                      # (X, Y, Z); tensors are (Z, Y, X, C)
    sensor: Sensor = field(default_factory=Sensor)
    # medium box: world = bbox_min + local * bbox_extent  (to_world = translate(-.5) scale(2))
    bbox_min: Tuple[float, float, float] = (-0.5, -0.5, -0.5)
    bbox_extent: Tuple[float, float, float] = (2.0, 2.0, 2.0)
    scale: float = 1.0                             # medium 'scale' (density_scale)
    majorant_resolution_factor: int = 0            # scene_config.py:36; <=1 disables supergrid
    radiance: Tuple[float, float, float] = (1.0, 0.8, 0.2)     # `constant` emitter ...
    envmap: "EnvMap" = None                                    # ... or an `envmap` emitter (then radiance is unused)

    def to_local(self) -> np.ndarray:
        m = np.zeros((3, 4), dtype=np.float64)
        for a in range(3):
            m[a, a] = 1.0 / self.bbox_extent[a]
            m[a, 3] = -self.bbox_min[a] / self.bbox_extent[a]
        return m.astype(np.float32)

    def effective_majorant_factor(self) -> int:
        """optimize.py:182-199 adjust_majorant_res_factor: shrink the factor until the
        supergrid has at least 4 cells per side; <=1 disables it."""
        f = int(self.majorant_resolution_factor)
        if f > 1:
            min_side = min(self.res)
            while f > 1 and (min_side // f) < 4:
                f -= 1
        return 0 if f <= 1 else f

    def with_sensor(self, sensor: Sensor) -> "VolumeScene":
        import copy
        s = copy.copy(self)
        s.sensor = sensor
        return s

    def with_resolution(self, res, majorant_resolution_factor: int) -> "VolumeScene":
        """The same scene with a re-sampled medium (multires upsampling, optimize.py:228-252)."""
        import copy
        s = copy.copy(self)
        s.res = tuple(int(r) for r in res)
        s.majorant_resolution_factor = int(majorant_resolution_factor)
        return s

    def as_dict(self) -> Dict[str, object]:
        d = {
            "res": tuple(int(r) for r in self.res),
            "to_local": self.to_local().reshape(-1),
            "scale": np.float32(self.scale),
            "majorant_factor": self.effective_majorant_factor(),
            "radiance": np.asarray(self.radiance, dtype=np.float32),
            # linear part of local -> world (directions of escaped rays are looked up in the envmap)
            "local_to_world": np.diag(np.asarray(self.bbox_extent, dtype=np.float64)).astype(np.float32).reshape(-1),
        }
        d.update(self.sensor.frame())
        if self.envmap is not None:
            d.update(self.envmap.tables())
        return d


# --------------------------------------------------------------------------------------
# Fixtures
# --------------------------------------------------------------------------------------

def cube_test_grids():
    """The 3x3x3 grids of tests/test_integrators.py:22-37 -> (sigma_t (3,3,3,1), albedo (3,3,3,3))."""
    sig = np.full((3, 3, 3, 1), 0.5, dtype=np.float32)
    sig[0, 0, 0, :] = 0.1
    sig[0, -1, 0, :] = 2.0
    sig[0, 0, -1, :] = 0.2
    g = np.full((3, 3, 3, 3), 1.0, dtype=np.float32)
    g[..., 0] = 0.3
    g[..., 1] = 0.5
    g[..., 2] = 0.9
    for i in range(3):
        g[i, :, :, 0] *= np.square((i + 1) / 3)
        g[i, :, :, 1] *= 1 - (i + 1) / 3
        g[:, i, :, 1] *= np.square((i + 1) / 3)
    return sig, np.clip(g, 0, 1).astype(np.float32)


def cube_test_scene(resx=128, resy=128, density_scale=1.0, res=(3, 3, 3)) -> VolumeScene:
    """Geometry of tests/test_integrators.py:19-116 (`cube_test_scene`)."""
    return VolumeScene(res=tuple(res), sensor=Sensor(width=resx, height=resy),
                       scale=density_scale, majorant_resolution_factor=0)


def synthetic_grids(n: int, seed: int = 20220721, dense: bool = False):
    """Deterministic heterogeneous sigma_t / albedo grids, SURVEY §8(d) recipe.
    Returns torch CPU tensors (sigma_t (n,n,n,1) in [0,1], albedo (n,n,n,3)).
    dense: a medium without empty space (no spherical fall-off, no `< 0.05 -> 0` cut; d = 0.5 + 0.5 f):
    the second, HBM-heavier bench workload."""
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
    return d.unsqueeze(-1).contiguous(), albedo


def benchmark_scene(n: int, width: int, height: int, scale: float = 8.0,
                    majorant_resolution_factor: int = 8) -> VolumeScene:
    """fixture geometry, camera aimed at the box centre."""
    return VolumeScene(res=(n, n, n),
                       sensor=Sensor(target=(0.5, 0.5, 0.5), width=width, height=height),
                       scale=scale, majorant_resolution_factor=majorant_resolution_factor)
