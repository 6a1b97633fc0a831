"""Coarse-to-fine pieces of the optimisation loop and `.vol` checkpoints -- rank-3 "next" row of
SURVEY §8(f).

Reference                                                    -> here
  upsample_grid            (optimize.py:203-225)               -> upsample_grid (CUDA, uivr_upsample2x)
  upsample_params_if_needed (optimize.py:228-252)              -> upsample_params
  adjust_majorant_res_factor (optimize.py:182-199)             -> adjust_majorant_res_factor
  OptimizationConfig.upsample_at / should_upsample (opt_config.py:40-44, :72-75) -> upsample_iterations
  save_params -> mi.VolumeGrid.write (util.py:55-71)           -> save_params / write_vol / read_vol
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterable, Sequence, Set, Tuple

import numpy as np
import torch

from . import _native
from .integrator import SIGMA_T_SUFFIX, Scene, _stream


def upsample_grid(ctx: _native.Context, values: torch.Tensor, new_res: Sequence[int]) -> torch.Tensor:
    """optimize.py:203-225: first-order `scipy.ndimage.zoom(..., mode='nearest', grid_mode=True)` of a
    (Z,Y,X,C) grid.  The reference only ever doubles the resolution (optimize.py:239); other factors
    are not implemented."""
    if values.dim() == 3:
        values = values[..., None]
    old_res = tuple(values.shape)
    new_res = tuple(int(r) for r in new_res)
    assert len(old_res) == 4 and len(new_res) == 4 and new_res[-1] == old_res[-1]
    if old_res == new_res:
        return values.detach().clone()
    if new_res[:3] != tuple(2 * r for r in old_res[:3]):
        raise NotImplementedError(f"only x2 upsampling is implemented: {old_res} -> {new_res}")
    src = values.detach().to(torch.float32).contiguous()
    out = torch.empty(new_res, dtype=torch.float32, device=src.device)
    z, y, x, c = old_res
    ctx.upsample2x(src.data_ptr(), (x, y, z), c, out.data_ptr(), _stream())
    return out


def adjust_majorant_res_factor(res_factor: int, density_res: Sequence[int]) -> int:
    """optimize.py:182-199: the largest factor <= res_factor that leaves a supergrid of at least 4
    cells along the shortest axis; 0 (supergrid off) when that is <= 1."""
    if res_factor > 1:
        min_side = min(int(r) for r in density_res[:3])
        while res_factor > 1 and (min_side // res_factor) < 4:
            res_factor -= 1
    if res_factor <= 1:
        res_factor = 0
    return res_factor


def upsample_iterations(upsample: Iterable[float], n_iter: int) -> Set[int]:
    """opt_config.py:40-44: iterations at which the grids are upsampled (fractions of the run)."""
    out = set()
    for t in upsample or ():
        assert 0 <= t <= 1
        out.add(int(t * n_iter))
    return out


def upsample_params(scene: Scene, opt, majorant_resolution_factor: int, keep_adjusted: bool = False) -> Dict[str, Tuple[int, ...]]:
    """optimize.py:228-252: double the resolution of every optimised grid and rebuild the medium.  `opt` is an
    optimize.Adam: the state of a re-shaped parameter starts over, as mi.ad.Optimizer does when a parameter changes
    size.  Returns the new shapes.

    Supergrid factor: the reference re-derives it for the new density resolution inside the loop (:245-246) and then
    sets the medium back to the CONFIGURED factor (:249-250), so after an upsampling step the configured factor is
    what renders -- mirrored here, because the tentative collisions (and with them the sample streams) depend on it.
    `keep_adjusted=True` keeps the re-derived factor instead (a supergrid of at least 4 cells per side)."""
    new_shapes = {}
    adjusted = majorant_resolution_factor
    k_sig = None
    for k in list(opt.params.keys()):
        v = opt.params[k]
        old_res = tuple(v.shape)
        assert len(old_res) == 4
        new_res = (*[2 * r for r in old_res[:3]], old_res[-1])
        up = upsample_grid(scene.ctx, v, new_res)
        opt.replace(k, up)
        new_shapes[k] = new_res
        if k.endswith(SIGMA_T_SUFFIX):
            k_sig = k
            adjusted = adjust_majorant_res_factor(majorant_resolution_factor, new_res)
    z, y, x = new_shapes[k_sig][:3]
    scene.volume = scene.volume.with_resolution((x, y, z), adjusted if keep_adjusted else majorant_resolution_factor)
    scene.update_medium_after_reshape(opt.params[k_sig])
    return new_shapes


# ---- Mitsuba 3 `.vol` grids (VolumeGrid::write): 'VOL' 0x03, int32 type (1 = float32), int32 x y z
# ---- channels, float32 bbox min xyz / max xyz, then float32 data, index ((z * Y + y) * X + x) * C + c

def write_vol(path: str, grid, bbox_min=(0.0, 0.0, 0.0), bbox_max=(1.0, 1.0, 1.0)) -> None:
    a = grid.detach().cpu().numpy() if isinstance(grid, torch.Tensor) else np.asarray(grid)
    if a.ndim == 3:
        a = a[..., None]
    assert a.ndim == 4, "expected a (Z,Y,X,C) grid"
    z, y, x, c = a.shape
    with open(path, "wb") as f:
        f.write(b"VOL\x03")
        f.write(struct.pack("<iiiii", 1, x, y, z, c))
        f.write(struct.pack("<6f", *[float(v) for v in bbox_min], *[float(v) for v in bbox_max]))
        f.write(np.ascontiguousarray(a, dtype="<f4").tobytes())


def read_vol(path: str):
    with open(path, "rb") as f:
        head = f.read(48)
        if len(head) != 48 or head[:3] != b"VOL" or head[3] != 3:
            raise ValueError(f"{path}: not a Mitsuba VOL v3 file")
        dtype, x, y, z, c = struct.unpack("<iiiii", head[4:24])
        if dtype != 1:
            raise NotImplementedError(f"{path}: only float32 grids (type 1) are supported, got {dtype}")
        bbox = struct.unpack("<6f", head[24:48])
        data = np.frombuffer(f.read(), dtype="<f4")
    if data.size != x * y * z * c:
        raise ValueError(f"{path}: expected {x * y * z * c} values, found {data.size}")
    return data.reshape(z, y, x, c).copy(), bbox[:3], bbox[3:]


def save_params(output_dir: str, params: Dict[str, torch.Tensor], name: str, keys: Iterable[str] = None) -> Dict[str, str]:
    """util.save_params (util.py:55-71): one `<name>-<var>.vol` per `*.data` parameter."""
    os.makedirs(output_dir, exist_ok=True)
    written = {}
    for key in (keys or params.keys()):
        if not key.endswith(".data"):
            raise NotImplementedError(f"Checkpointing of parameter {key} with type {type(params[key])}")
        var_name = "_".join(key[:-len(".data")].strip().split("."))
        fname = os.path.join(output_dir, f"{name}-{var_name}.vol")
        write_vol(fname, params[key])
        written[key] = fname
    return written
