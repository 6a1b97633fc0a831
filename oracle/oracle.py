"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE -- not the product).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Pinned against the reference's own Python files through oracle/refshim.py
(tests/golden/refshim_*.npz); PARITY UNPINNED only for the arithmetic inside the un-vendored
Mitsuba 3 branch: see oracle/uivr_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

COUNTER_NAMES = ["sigma_taps", "albedo_taps", "majorant_reads", "sigma_scatters",
                 "albedo_scatters", "camera_hits", "real_collisions", "rng_draws", "samples"]


class _Scene(C.Structure):
    _fields_ = [
        ("res", C.c_int32 * 3), ("to_local", C.c_float * 12), ("scale", C.c_float),
        ("majorant_factor", C.c_int32),
        ("cam_origin", C.c_float * 3), ("cam_left", C.c_float * 3), ("cam_up", C.c_float * 3),
        ("cam_dir", C.c_float * 3), ("tan_x", C.c_float), ("tan_y", C.c_float),
        ("near_clip", C.c_float), ("width", C.c_int32), ("height", C.c_int32),
        ("radiance", C.c_float * 3),
        ("max_depth", C.c_int32), ("hide_emitters", C.c_int32), ("use_nee", C.c_int32),
        ("use_drt", C.c_int32), ("use_drt_subsampling", C.c_int32), ("use_drt_mis", C.c_int32),
        ("local_to_world", C.c_float * 9), ("env_data", C.POINTER(C.c_float)), ("env_w", C.c_int32),
        ("env_h", C.c_int32), ("env_scale", C.c_float), ("env_marg", C.POINTER(C.c_float)),
        ("env_cond", C.POINTER(C.c_float)), ("env_to_world", C.c_float * 9), ("world_to_env", C.c_float * 9),
    ]


class _Shard(C.Structure):
    _fields_ = [("shard_rank", C.c_int32), ("shard_count", C.c_int32), ("shard_block", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, oracle/Makefile) -> oracle/_build/*.so."""
    so = os.path.join(_BUILD, "libuivr_oracle.so")
    src = os.path.join(_HERE, "uivr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            return " fma " in f.read().replace("\n", " ")
    except OSError:
        return False


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    name = "libuivr_oracle.so" if _has_fma() else "libuivr_oracle_nofma.so"
    path = os.path.join(_BUILD, name)
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    fp, dp, u64p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    L.uivr_oracle_render_forward.argtypes = [C.POINTER(_Scene), fp, fp, C.c_uint32, C.c_int32,
                                             C.POINTER(_Shard), C.c_int, fp, fp, u64p]
    L.uivr_oracle_render_forward.restype = C.c_int
    L.uivr_oracle_render_backward.argtypes = [C.POINTER(_Scene), fp, fp, fp, C.c_uint32, C.c_int32,
                                              C.POINTER(_Shard), C.c_int, dp, dp, fp, u64p]
    L.uivr_oracle_render_backward.restype = C.c_int
    L.uivr_oracle_alt_seed.argtypes = [C.c_uint32]
    L.uivr_oracle_alt_seed.restype = C.c_uint32
    L.uivr_oracle_alt_seed_batch.argtypes = [C.c_uint32]
    L.uivr_oracle_alt_seed_batch.restype = C.c_uint32
    _lib = L
    return L


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a: Optional[np.ndarray], t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def make_scene(desc: Dict, props: Dict) -> _Scene:
    """desc: VolumeScene.as_dict(); props: integrator properties (volpathsimple.py:19-34)."""
    s = _Scene()
    s.res[:] = [int(v) for v in desc["res"]]
    s.to_local[:] = [float(v) for v in np.asarray(desc["to_local"]).reshape(-1)]
    s.scale = float(desc["scale"])
    s.majorant_factor = int(desc["majorant_factor"])
    for k in ("cam_origin", "cam_left", "cam_up", "cam_dir", "radiance"):
        getattr(s, k)[:] = [float(v) for v in desc[k]]
    s.tan_x, s.tan_y, s.near_clip = float(desc["tan_x"]), float(desc["tan_y"]), float(desc["near_clip"])
    s.width, s.height = int(desc["width"]), int(desc["height"])
    s.max_depth = int(props["max_depth"])
    s.hide_emitters = int(bool(props.get("hide_emitters", False)))
    s.use_nee = int(bool(props.get("use_nee", True)))
    s.use_drt = int(bool(props.get("use_drt", True)))
    s.use_drt_subsampling = int(bool(props.get("use_drt_subsampling", True)))
    s.use_drt_mis = int(bool(props.get("use_drt_mis", True)))
    l2w = desc.get("local_to_world")
    if l2w is None:  # linear part of the inverse of to_local
        l2w = np.linalg.inv(np.asarray(desc["to_local"], dtype=np.float64).reshape(3, 4)[:, :3]).reshape(-1)
    s.local_to_world[:] = [float(v) for v in np.asarray(l2w).reshape(-1)]
    if desc.get("env_data") is not None:
        keep = [_f32(desc[k]) for k in ("env_data", "env_marg", "env_cond")]
        s._keepalive = keep  # the struct only holds raw pointers
        s.env_data, s.env_marg, s.env_cond = (_ptr(a, C.c_float) for a in keep)
        s.env_w, s.env_h, s.env_scale = int(desc["env_w"]), int(desc["env_h"]), float(desc["env_scale"])
        s.env_to_world[:] = [float(v) for v in desc["env_to_world"]]
        s.world_to_env[:] = [float(v) for v in desc["world_to_env"]]
    return s


def _shard(shard):
    if shard is None:
        return None
    s = _Shard()
    s.shard_rank, s.shard_count, s.shard_block = (int(v) for v in shard)
    return C.byref(s)


def _check_grids(desc, sigma_t, albedo):
    x, y, z = desc["res"]
    sigma_t = _f32(sigma_t).reshape(z, y, x)
    albedo = _f32(albedo).reshape(z, y, x, 3)
    return sigma_t, albedo


def render_forward(desc, props, sigma_t, albedo, seed: int, spp: int, shard=None,
                   nthreads: Optional[int] = None, want_samples: bool = False):
    """-> (image (H,W,3) f32, per-sample L (S,3) or None, counters dict)."""
    L = lib()
    sc = make_scene(desc, props)
    sigma_t, albedo = _check_grids(desc, sigma_t, albedo)
    h, w = sc.height, sc.width
    image = np.zeros((h, w, 3), dtype=np.float32)
    samples = np.zeros((h * w * spp, 3), dtype=np.float32) if want_samples else None
    counters = np.zeros(len(COUNTER_NAMES), dtype=np.uint64)
    rc = L.uivr_oracle_render_forward(C.byref(sc), _ptr(sigma_t, C.c_float), _ptr(albedo, C.c_float),
                                      seed & 0xFFFFFFFF, spp, _shard(shard),
                                      nthreads or os.cpu_count() or 1,
                                      _ptr(image, C.c_float), _ptr(samples, C.c_float),
                                      _ptr(counters, C.c_uint64))
    if rc != 0:
        raise RuntimeError(f"uivr_oracle_render_forward failed ({rc})")
    return image, samples, dict(zip(COUNTER_NAMES, (int(c) for c in counters)))


def render_backward(desc, props, sigma_t, albedo, grad_image, seed_grad: int, spp_grad: int,
                    shard=None, nthreads: Optional[int] = None, want_samples: bool = False):
    """-> (d sigma_t (Z,Y,X,1) f64, d albedo (Z,Y,X,3) f64, per-sample primal L or None, counters)."""
    L = lib()
    sc = make_scene(desc, props)
    sigma_t, albedo = _check_grids(desc, sigma_t, albedo)
    x, y, z = desc["res"]
    h, w = sc.height, sc.width
    grad_image = _f32(grad_image).reshape(h, w, 3)
    dsig = np.zeros((z, y, x, 1), dtype=np.float64)
    dalb = np.zeros((z, y, x, 3), dtype=np.float64)
    samples = np.zeros((h * w * spp_grad, 3), dtype=np.float32) if want_samples else None
    counters = np.zeros(len(COUNTER_NAMES), dtype=np.uint64)
    rc = L.uivr_oracle_render_backward(C.byref(sc), _ptr(sigma_t, C.c_float), _ptr(albedo, C.c_float),
                                       _ptr(grad_image, C.c_float), seed_grad & 0xFFFFFFFF, spp_grad,
                                       _shard(shard), nthreads or os.cpu_count() or 1,
                                       _ptr(dsig, C.c_double), _ptr(dalb, C.c_double),
                                       _ptr(samples, C.c_float), _ptr(counters, C.c_uint64))
    if rc != 0:
        raise RuntimeError(f"uivr_oracle_render_backward failed ({rc})")
    return dsig, dalb, samples, dict(zip(COUNTER_NAMES, (int(c) for c in counters)))


def last_backward_primal_counters() -> Dict[str, int]:
    """Events of the primal pass (batched.py:255-264) inside the most recent backward call of this process; they
    are included in that call's counters."""
    out = np.zeros(len(COUNTER_NAMES), dtype=np.uint64)
    lib().uivr_oracle_last_backward_primal_counters(_ptr(out, C.c_uint64))
    return dict(zip(COUNTER_NAMES, (int(c) for c in out)))


def last_backward_replay_counters() -> Dict[str, int]:
    """Events of the NEE adjoint's second walk over the shadow segments (volpathsimple.py:393-401) inside the most
    recent backward call."""
    out = np.zeros(len(COUNTER_NAMES), dtype=np.uint64)
    lib().uivr_oracle_last_backward_replay_counters(_ptr(out, C.c_uint64))
    return dict(zip(COUNTER_NAMES, (int(c) for c in out)))


def fused_backward_counters(total: Dict[str, int]) -> Dict[str, int]:
    """Event counts of the CUDA slot-pool backward, from the counts of the backward just computed here: its adjoint
    replay gathers the primal radiance itself (no primal pass: those events are not executed; samples and camera
    hits are counted once) and logs the tentative collisions of every NEE shadow walk instead of walking the
    segment a second time for the adjoint (no second set of supergrid reads / taps / draws; the scatters stay)."""
    primal = last_backward_primal_counters()
    replay = last_backward_replay_counters()
    out = {k: total[k] - primal[k] for k in total}
    out["camera_hits"] = total["camera_hits"]
    out["samples"] = total["samples"]
    for k in ("sigma_taps", "majorant_reads", "rng_draws"):
        out[k] -= replay[k]
    return out


# ---- primitives ----

def tea(v0: int, v1: int):
    out = (C.c_uint32 * 2)()
    lib().uivr_oracle_tea(C.c_uint32(v0 & 0xFFFFFFFF), C.c_uint32(v1 & 0xFFFFFFFF), out)
    return int(out[0]), int(out[1])


def pcg32_stream(initstate: int, initseq: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint32)
    lib().uivr_oracle_pcg32_stream(C.c_uint64(initstate), C.c_uint64(initseq), n, _ptr(out, C.c_uint32))
    return out


def sampler_floats(seed: int, idx: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.float32)
    lib().uivr_oracle_sampler_floats(C.c_uint32(seed & 0xFFFFFFFF), C.c_uint32(idx), n, _ptr(out, C.c_float))
    return out


def neg_log1m(u) -> np.ndarray:
    u = _f32(u)
    out = np.zeros_like(u)
    lib().uivr_oracle_neg_log1m(_ptr(u, C.c_float), u.size, _ptr(out, C.c_float))
    return out


def sincos2pi(x):
    x = _f32(x)
    s, c = np.zeros_like(x), np.zeros_like(x)
    lib().uivr_oracle_sincos2pi(_ptr(x, C.c_float), x.size, _ptr(s, C.c_float), _ptr(c, C.c_float))
    return s, c


def alt_seed(seed_grad: int) -> int:
    return int(lib().uivr_oracle_alt_seed(seed_grad & 0xFFFFFFFF))


def alt_seed_batch(seed_grad: int) -> int:
    return int(lib().uivr_oracle_alt_seed_batch(seed_grad & 0xFFFFFFFF))


def trilinear(grid, p) -> np.ndarray:
    grid = _f32(grid)
    z, y, x, ch = grid.shape
    p = _f32(p).reshape(-1, 3)
    out = np.zeros((p.shape[0], ch), dtype=np.float32)
    res = (C.c_int32 * 3)(x, y, z)
    lib().uivr_oracle_trilinear(_ptr(grid, C.c_float), res, ch, _ptr(p, C.c_float), p.shape[0], _ptr(out, C.c_float))
    return out


def build_majorant(sigma_t, scale: float, factor: int) -> np.ndarray:
    sigma_t = _f32(sigma_t)
    z, y, x = sigma_t.shape[:3]
    res = (C.c_int32 * 3)(x, y, z)
    mres = (C.c_int32 * 3)()
    m = [max(1, r // factor) if factor > 1 else 1 for r in (x, y, z)]
    out = np.zeros((m[2], m[1], m[0]), dtype=np.float32)
    lib().uivr_oracle_build_majorant(_ptr(sigma_t, C.c_float), res, C.c_float(scale), factor, mres, _ptr(out, C.c_float))
    assert list(mres) == m
    return out


def build_exit_mask(majorant) -> np.ndarray:
    """(MZ,MY,MX) uint8: bit o of a cell = only empty supergrid cells ahead in octant o (uivr_oracle.c)."""
    majorant = _f32(majorant)
    mz, my, mx = majorant.shape
    out = np.zeros((mz, my, mx), dtype=np.uint8)
    lib().uivr_oracle_build_exit_mask(_ptr(majorant, C.c_float), (C.c_int32 * 3)(mx, my, mz), _ptr(out, C.c_uint8))
    return out


def set_exit_mask(enable: bool) -> None:
    """Test hook: walks stop early in empty space (default) or always run to the medium boundary."""
    lib().uivr_oracle_set_exit_mask(1 if enable else 0)


def set_nee_log_capacity(capacity: int) -> None:
    """Counters only: model a collision log of `capacity` entries per NEE shadow walk (the CUDA adjoint kernel's
    kNeeLog): walks with more tentative collisions are walked twice there as well.  0 = unlimited (default)."""
    lib().uivr_oracle_set_nee_log_capacity(int(capacity))


def nee_log_overflows() -> int:
    f = lib().uivr_oracle_nee_log_overflows
    f.restype = C.c_uint64
    return int(f())


def set_remaining_by_difference(enable: bool) -> None:
    """Test hook: Li of the adjoint as L - gathered (the CUDA pipeline's form) instead of the running subtraction."""
    lib().uivr_oracle_set_remaining_by_difference(1 if enable else 0)


def adam_step(param, grad, m, v, lr, beta1, beta2, eps, t, lo, hi):
    """In place on float32 arrays: mi.ad.Adam step + clip (optimize.py:169-179, :352-353)."""
    n = param.size
    for a in (param, grad, m, v):
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"] and a.size == n
    lib().uivr_oracle_adam_step(_ptr(param, C.c_float), _ptr(grad, C.c_float), _ptr(m, C.c_float), _ptr(v, C.c_float),
                                C.c_uint64(n), C.c_float(lr), C.c_float(beta1), C.c_float(beta2), C.c_float(eps),
                                C.c_int32(t), C.c_float(lo), C.c_float(hi))


# ---- ray-batch rendering (python/batched.py) ----

class _Batch(C.Structure):
    _fields_ = [("n_sensors", C.c_int32), ("sensors", C.POINTER(C.c_float)), ("film_w", C.c_int32),
                ("film_h", C.c_int32), ("seed_pixels", C.c_uint32), ("seed_offsets", C.c_uint32)]


def _make_batch(sensors16, film_size, seed_pixels, seed_offsets):
    a = np.ascontiguousarray(sensors16, dtype=np.float32).reshape(-1, 16)
    b = _Batch()
    b.n_sensors = a.shape[0]
    b.sensors = a.ctypes.data_as(C.POINTER(C.c_float))
    b.film_w, b.film_h = int(film_size[0]), int(film_size[1])
    b.seed_pixels, b.seed_offsets = seed_pixels & 0xFFFFFFFF, seed_offsets & 0xFFFFFFFF
    return b, a  # keep `a` alive


def batch_elements(sensors16, film_size, seed: int, batch_size: int) -> np.ndarray:
    """[B, 3] (sensor, px, py) per batched.py:397-423 for render_batch(seed=seed)."""
    b, keep = _make_batch(sensors16, film_size, tea(seed, 5)[0], 0)
    out = np.zeros((batch_size, 3), dtype=np.uint32)
    e = (C.c_uint32 * 3)()
    for i in range(batch_size):
        lib().uivr_oracle_batch_element(C.byref(b), C.c_uint32(i), e)
        out[i] = (e[0], e[1], e[2])
    return out


def render_batch_forward(desc, props, sensors16, film_size, batch_size, sigma_t, albedo, seed, spp,
                         nthreads: Optional[int] = None, want_samples: bool = False):
    """batched.py render_batch primal: image [B, 3]; desc supplies the medium / emitter only."""
    d = dict(desc)
    d["width"], d["height"] = int(batch_size), 1
    sc = make_scene(d, props)
    sigma_t, albedo = _check_grids(desc, sigma_t, albedo)
    b, keep = _make_batch(sensors16, film_size, tea(seed, 5)[0], tea(seed, 22)[0])
    image = np.zeros((batch_size, 3), dtype=np.float32)
    samples = np.zeros((batch_size * spp, 3), dtype=np.float32) if want_samples else None
    counters = np.zeros(len(COUNTER_NAMES), dtype=np.uint64)
    rc = lib().uivr_oracle_render_batch_forward(C.byref(sc), C.byref(b), _ptr(sigma_t, C.c_float), _ptr(albedo, C.c_float),
                                                seed & 0xFFFFFFFF, spp, nthreads or os.cpu_count() or 1,
                                                _ptr(image, C.c_float), _ptr(samples, C.c_float), _ptr(counters, C.c_uint64))
    if rc != 0:
        raise RuntimeError(f"uivr_oracle_render_batch_forward failed ({rc})")
    return image, samples, dict(zip(COUNTER_NAMES, (int(c) for c in counters)))


def render_batch_backward(desc, props, sensors16, film_size, batch_size, sigma_t, albedo, grad_image, seed, seed_grad,
                          spp_grad, nthreads: Optional[int] = None, want_samples: bool = False):
    """batched.py render_batch adjoint: `seed` selects the pixels, `seed_grad` the paths."""
    d = dict(desc)
    d["width"], d["height"] = int(batch_size), 1
    sc = make_scene(d, props)
    sigma_t, albedo = _check_grids(desc, sigma_t, albedo)
    x, y, z = desc["res"]
    b, keep = _make_batch(sensors16, film_size, tea(seed, 5)[0], tea(seed, 39)[0])
    grad_image = _f32(grad_image).reshape(batch_size, 3)
    dsig = np.zeros((z, y, x, 1), dtype=np.float64)
    dalb = np.zeros((z, y, x, 3), dtype=np.float64)
    samples = np.zeros((batch_size * spp_grad, 3), dtype=np.float32) if want_samples else None
    counters = np.zeros(len(COUNTER_NAMES), dtype=np.uint64)
    rc = lib().uivr_oracle_render_batch_backward(C.byref(sc), C.byref(b), _ptr(sigma_t, C.c_float), _ptr(albedo, C.c_float),
                                                 _ptr(grad_image, C.c_float), seed_grad & 0xFFFFFFFF, spp_grad,
                                                 nthreads or os.cpu_count() or 1, _ptr(dsig, C.c_double),
                                                 _ptr(dalb, C.c_double), _ptr(samples, C.c_float), _ptr(counters, C.c_uint64))
    if rc != 0:
        raise RuntimeError(f"uivr_oracle_render_batch_backward failed ({rc})")
    return dsig, dalb, samples, dict(zip(COUNTER_NAMES, (int(c) for c in counters)))


# ---- nerf integrator (python/integrators/nerf.py) ----

class _Nerf(C.Structure):
    _fields_ = [("queries_per_ray", C.c_int32), ("jittering_enabled", C.c_int32), ("activation", C.c_int32),
                ("hide_emitters", C.c_int32)]


def make_nerf(props: Dict) -> _Nerf:
    """props: NeRFIntegrator properties (nerf.py:27-35)."""
    n = _Nerf()
    n.queries_per_ray = int(props.get("queries_per_ray", 128))
    n.jittering_enabled = int(bool(props.get("jittering_enabled", True)))
    act = str(props.get("activation", "identity")).lower()
    if act not in ("identity", "relu"):
        raise ValueError(f"Unsupported activation: {act}")  # nerf.py:44
    n.activation = 1 if act == "relu" else 0
    n.hide_emitters = int(bool(props.get("hide_emitters", False)))
    return n


def _nerf_batch(desc, batch, seed, i_offsets):
    """batch = (sensors16, film_size, batch_size, seed of render_batch) or None"""
    if batch is None:
        return desc, None, None
    sensors16, film_size, batch_size, bseed = batch
    d = dict(desc)
    d["width"], d["height"] = int(batch_size), 1
    b, keep = _make_batch(sensors16, film_size, tea(bseed, 5)[0], tea(bseed, i_offsets)[0])
    return d, b, keep


def nerf_forward(desc, props, sigma_t, emission, seed: int, spp: int, shard=None,
                 nthreads: Optional[int] = None, want_samples: bool = False, batch=None):
    """-> (image (H,W,3) [or (B,3) in ray-batch mode] f32, per-sample L (S,3) or None, counters dict)."""
    sigma_t, emission = _check_grids(desc, sigma_t, emission)
    desc, b, keep = _nerf_batch(desc, batch, seed, 22)
    sc = make_scene(desc, dict(max_depth=0))
    nf = make_nerf(props)
    h, w = sc.height, sc.width
    image = np.zeros((h, w, 3) if batch is None else (w, 3), dtype=np.float32)
    samples = np.zeros((h * w * spp, 3), dtype=np.float32) if want_samples else None
    counters = np.zeros(len(COUNTER_NAMES), dtype=np.uint64)
    rc = lib().uivr_oracle_nerf_forward(C.byref(sc), C.byref(nf), None if b is None else C.byref(b),
                                        _ptr(sigma_t, C.c_float), _ptr(emission, C.c_float),
                                        C.c_uint32(seed & 0xFFFFFFFF), C.c_int32(spp), _shard(shard),
                                        C.c_int(nthreads or os.cpu_count() or 1), _ptr(image, C.c_float),
                                        _ptr(samples, C.c_float), _ptr(counters, C.c_uint64))
    if rc != 0:
        raise RuntimeError(f"uivr_oracle_nerf_forward failed ({rc})")
    return image, samples, dict(zip(COUNTER_NAMES, (int(c) for c in counters)))


def nerf_backward(desc, props, sigma_t, emission, grad_image, seed_grad: int, spp_grad: int, shard=None,
                  nthreads: Optional[int] = None, want_samples: bool = False, batch=None):
    """-> (d sigma_t (Z,Y,X,1) f64, d emission (Z,Y,X,3) f64, per-sample primal L or None, counters)."""
    sigma_t, emission = _check_grids(desc, sigma_t, emission)
    x, y, z = desc["res"]
    desc, b, keep = _nerf_batch(desc, batch, seed_grad, 39)
    sc = make_scene(desc, dict(max_depth=0))
    nf = make_nerf(props)
    h, w = sc.height, sc.width
    grad_image = _f32(grad_image).reshape(h, w, 3)
    dsig = np.zeros((z, y, x, 1), dtype=np.float64)
    dem = np.zeros((z, y, x, 3), dtype=np.float64)
    samples = np.zeros((h * w * spp_grad, 3), dtype=np.float32) if want_samples else None
    counters = np.zeros(len(COUNTER_NAMES), dtype=np.uint64)
    rc = lib().uivr_oracle_nerf_backward(C.byref(sc), C.byref(nf), None if b is None else C.byref(b),
                                         _ptr(sigma_t, C.c_float), _ptr(emission, C.c_float),
                                         _ptr(grad_image, C.c_float), C.c_uint32(seed_grad & 0xFFFFFFFF),
                                         C.c_int32(spp_grad), _shard(shard), C.c_int(nthreads or os.cpu_count() or 1),
                                         _ptr(dsig, C.c_double), _ptr(dem, C.c_double), _ptr(samples, C.c_float),
                                         _ptr(counters, C.c_uint64))
    if rc != 0:
        raise RuntimeError(f"uivr_oracle_nerf_backward failed ({rc})")
    return dsig, dem, samples, dict(zip(COUNTER_NAMES, (int(c) for c in counters)))


def exp_exact(x) -> np.ndarray:
    x = _f32(x)
    out = np.zeros_like(x)
    lib().uivr_oracle_exp(_ptr(x, C.c_float), x.size, _ptr(out, C.c_float))
    return out
