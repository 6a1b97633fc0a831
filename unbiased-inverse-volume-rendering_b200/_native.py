"""ctypes binding of libuivr.so (include/uivr.h).  There is NO fallback: if the CUDA library
is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
# UIVR_LIB: alternative build of the same library (tuning sweeps); default = the in-tree build
LIB_PATH = os.environ.get("UIVR_LIB") or os.path.join(_CSRC, "libuivr.so")

COUNTER_NAMES = ["sigma_taps", "albedo_taps", "majorant_reads", "sigma_scatters",
                 "albedo_scatters", "camera_hits", "real_collisions", "rng_draws", "samples"]

# algorithmic bytes per event
COUNTER_BYTES = {"sigma_taps": 32, "albedo_taps": 96, "majorant_reads": 4,
                 "sigma_scatters": 64, "albedo_scatters": 192}

EXPORTS = [
    "uivr_version", "uivr_create", "uivr_destroy", "uivr_last_error", "uivr_set_scene",
    "uivr_set_integrator", "uivr_update_medium", "uivr_render_forward", "uivr_render_backward",
    "uivr_render_forward_host", "uivr_render_backward_host", "uivr_set_counting",
    "uivr_reset_counters", "uivr_get_counters", "uivr_get_kernel_ms", "uivr_get_launch_count",
    "uivr_set_variant", "uivr_debug_set_walk_limit", "uivr_check_watchdog", "uivr_adam_step", "uivr_set_batch", "uivr_upsample2x",
    "uivr_test_neg_log1m", "uivr_test_sincos2pi", "uivr_test_sampler", "uivr_test_sigma_lookup",
    "uivr_get_majorant", "uivr_get_walk_table", "uivr_tea32", "uivr_alt_seed", "uivr_alt_seed_batch",
    "uivr_nerf_forward", "uivr_nerf_backward", "uivr_test_exp", "uivr_set_envmap", "uivr_test_atan2_turns",
]


class NativeError(RuntimeError):
    pass


class SceneDesc(C.Structure):
    _fields_ = [
        ("res", C.c_int32 * 3), ("to_local", C.c_float * 12), ("scale", C.c_float),
        ("majorant_factor", C.c_int32),
        ("cam_origin", C.c_float * 3), ("cam_left", C.c_float * 3), ("cam_up", C.c_float * 3),
        ("cam_dir", C.c_float * 3), ("tan_x", C.c_float), ("tan_y", C.c_float),
        ("near_clip", C.c_float), ("width", C.c_int32), ("height", C.c_int32),
        ("radiance", C.c_float * 3),
    ]


class IntegratorProps(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("hide_emitters", C.c_int32), ("use_nee", C.c_int32),
                ("use_drt", C.c_int32), ("use_drt_subsampling", C.c_int32), ("use_drt_mis", C.c_int32)]


class BatchDesc(C.Structure):
    _fields_ = [("n_sensors", C.c_int32), ("sensors", C.POINTER(C.c_float)), ("film_w", C.c_int32),
                ("film_h", C.c_int32), ("batch_size", C.c_int32), ("seed", C.c_uint32)]


class NerfProps(C.Structure):
    _fields_ = [("queries_per_ray", C.c_int32), ("jittering_enabled", C.c_int32), ("activation", C.c_int32),
                ("hide_emitters", C.c_int32)]


class EnvMapDesc(C.Structure):
    _fields_ = [("env_w", C.c_int32), ("env_h", C.c_int32), ("scale", C.c_float), ("data", C.POINTER(C.c_float)),
                ("marg", C.POINTER(C.c_float)), ("cond", C.POINTER(C.c_float)), ("env_to_world", C.c_float * 9),
                ("world_to_env", C.c_float * 9), ("local_to_world", C.c_float * 9)]


class Shard(C.Structure):
    _fields_ = [("rank", C.c_int32), ("count", C.c_int32), ("block", C.c_int32)]


def build(verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> csrc/libuivr.so (in-tree)."""
    env = dict(os.environ)
    if verbose:
        env["UIVR_NVCC_EXTRA"] = "-Xptxas -v"
    subprocess.check_call(["bash", os.path.join(_CSRC, "build.sh")], env=env)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(f"{LIB_PATH} is missing: run __graft_entry__.build() "
                          f"(there is no CPU / PyTorch fallback for the render path)")
    L = C.CDLL(LIB_PATH)
    vp, fp, u32, i32 = C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32
    sig = {
        "uivr_version": ([], C.c_int),
        "uivr_create": ([C.c_int, C.POINTER(vp)], C.c_int),
        "uivr_destroy": ([vp], C.c_int),
        "uivr_last_error": ([vp], C.c_char_p),
        "uivr_set_scene": ([vp, C.POINTER(SceneDesc)], C.c_int),
        "uivr_set_integrator": ([vp, C.POINTER(IntegratorProps)], C.c_int),
        "uivr_set_batch": ([vp, C.POINTER(BatchDesc)], C.c_int),
        "uivr_upsample2x": ([vp, fp, C.POINTER(C.c_int32), i32, fp, vp], C.c_int),
        "uivr_update_medium": ([vp, fp, vp], C.c_int),
        "uivr_render_forward": ([vp, fp, u32, i32, C.POINTER(Shard), fp, fp, vp], C.c_int),
        "uivr_render_backward": ([vp, fp, fp, u32, i32, C.POINTER(Shard), fp, fp, fp, vp], C.c_int),
        "uivr_render_forward_host": ([vp, fp, fp, u32, i32, C.POINTER(Shard), fp, vp], C.c_int),
        "uivr_render_backward_host": ([vp, fp, fp, fp, u32, i32, C.POINTER(Shard), fp, fp, vp], C.c_int),
        "uivr_set_counting": ([vp, C.c_int], C.c_int),
        "uivr_reset_counters": ([vp, vp], C.c_int),
        "uivr_get_counters": ([vp, C.POINTER(C.c_uint64), vp], C.c_int),
        "uivr_get_kernel_ms": ([vp, C.c_int, C.POINTER(C.c_float)], C.c_int),
        "uivr_get_launch_count": ([vp, C.POINTER(C.c_uint64)], C.c_int),
        "uivr_set_variant": ([vp, C.c_int], C.c_int),
        "uivr_debug_set_walk_limit": ([vp, C.c_int], C.c_int),
        "uivr_check_watchdog": ([vp, C.POINTER(C.c_uint32), vp], C.c_int),
        "uivr_adam_step": ([vp, fp, fp, fp, fp, C.c_uint64, C.c_float, C.c_float, C.c_float, C.c_float, i32,
                            C.c_float, C.c_float, vp], C.c_int),
        "uivr_test_neg_log1m": ([vp, fp, C.c_int, fp, vp], C.c_int),
        "uivr_test_sincos2pi": ([vp, fp, C.c_int, fp, fp, vp], C.c_int),
        "uivr_test_sampler": ([vp, u32, u32, C.c_int, C.c_int, fp, vp], C.c_int),
        "uivr_test_sigma_lookup": ([vp, fp, C.c_int, fp, vp], C.c_int),
        "uivr_get_majorant": ([vp, C.POINTER(C.c_int32), fp, vp], C.c_int),
        "uivr_get_walk_table": ([vp, C.POINTER(C.c_int32), fp, vp], C.c_int),
        "uivr_tea32": ([u32, u32], u32),
        "uivr_alt_seed": ([u32], u32),
        "uivr_alt_seed_batch": ([u32], u32),
        "uivr_nerf_forward": ([vp, C.POINTER(NerfProps), fp, u32, i32, C.POINTER(Shard), fp, fp, vp], C.c_int),
        "uivr_nerf_backward": ([vp, C.POINTER(NerfProps), fp, fp, u32, i32, C.POINTER(Shard), fp, fp, fp, vp], C.c_int),
        "uivr_test_exp": ([vp, fp, C.c_int, fp, vp], C.c_int),
        "uivr_set_envmap": ([vp, C.POINTER(EnvMapDesc)], C.c_int),
        "uivr_test_atan2_turns": ([vp, fp, fp, C.c_int, fp, vp], C.c_int),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = L
    return L


def tea32(v0: int, v1: int) -> int:
    """mi.sample_tea_32(v0, v1)[0] (batched.py:121; optimize.py:327-328)."""
    return int(lib().uivr_tea32(v0 & 0xFFFFFFFF, v1 & 0xFFFFFFFF))


class Context:
    """One uivr_ctx: bound to one CUDA device, not thread-safe."""

    def __init__(self, device: int = 0):
        self._L = lib()
        self._h = C.c_void_p()
        rc = self._L.uivr_create(int(device), C.byref(self._h))
        if rc != 0:
            raise NativeError(f"uivr_create(device={device}) failed with status {rc} "
                              f"(no CUDA device? there is no CPU fallback)")
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.uivr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self._L.uivr_last_error(self._h)
            raise NativeError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    # -- configuration --
    def set_scene(self, desc: dict):
        s = SceneDesc()
        s.res[:] = [int(v) for v in desc["res"]]
        s.to_local[:] = [float(v) for v in desc["to_local"]]
        s.scale = float(desc["scale"])
        s.majorant_factor = int(desc["majorant_factor"])
        for k in ("cam_origin", "cam_left", "cam_up", "cam_dir", "radiance"):
            getattr(s, k)[:] = [float(v) for v in desc[k]]
        s.tan_x, s.tan_y, s.near_clip = float(desc["tan_x"]), float(desc["tan_y"]), float(desc["near_clip"])
        s.width, s.height = int(desc["width"]), int(desc["height"])
        self._check(self._L.uivr_set_scene(self._h, C.byref(s)), "uivr_set_scene")

    def set_integrator(self, props: dict):
        p = IntegratorProps()
        p.max_depth = int(props["max_depth"])
        p.hide_emitters = int(bool(props.get("hide_emitters", False)))
        p.use_nee = int(bool(props.get("use_nee", True)))
        p.use_drt = int(bool(props.get("use_drt", True)))
        p.use_drt_subsampling = int(bool(props.get("use_drt_subsampling", True)))
        p.use_drt_mis = int(bool(props.get("use_drt_mis", True)))
        self._check(self._L.uivr_set_integrator(self._h, C.byref(p)), "uivr_set_integrator")

    def set_batch(self, sensors16=None, film_w: int = 0, film_h: int = 0, batch_size: int = 0, seed: int = 0):
        """Enter ray-batch mode (sensors16: float32 array [n, 16]) or leave it (sensors16 = None)."""
        if sensors16 is None:
            self._check(self._L.uivr_set_batch(self._h, None), "uivr_set_batch")
            return
        import numpy as np
        a = np.ascontiguousarray(sensors16, dtype=np.float32).reshape(-1, 16)
        d = BatchDesc()
        d.n_sensors = a.shape[0]
        d.sensors = a.ctypes.data_as(C.POINTER(C.c_float))
        d.film_w, d.film_h, d.batch_size, d.seed = int(film_w), int(film_h), int(batch_size), seed & 0xFFFFFFFF
        self._check(self._L.uivr_set_batch(self._h, C.byref(d)), "uivr_set_batch")

    def set_envmap(self, desc: Optional[dict]):
        """Switch the emitter to the envmap described by `desc` (the env_* / *_to_* entries of
        VolumeScene.as_dict()) or back to the constant emitter (desc without env_data / None)."""
        if desc is None or desc.get("env_data") is None:
            self._check(self._L.uivr_set_envmap(self._h, None), "uivr_set_envmap")
            return
        import numpy as np
        keep = [np.ascontiguousarray(desc[k], dtype=np.float32) for k in ("env_data", "env_marg", "env_cond")]
        e = EnvMapDesc()
        e.env_w, e.env_h, e.scale = int(desc["env_w"]), int(desc["env_h"]), float(desc["env_scale"])
        e.data, e.marg, e.cond = (a.ctypes.data_as(C.POINTER(C.c_float)) for a in keep)
        for k in ("env_to_world", "world_to_env", "local_to_world"):
            getattr(e, k)[:] = [float(v) for v in np.asarray(desc[k]).reshape(-1)]
        self._check(self._L.uivr_set_envmap(self._h, C.byref(e)), "uivr_set_envmap")

    def set_variant(self, variant: int):
        self._check(self._L.uivr_set_variant(self._h, int(variant)), "uivr_set_variant")

    def debug_set_walk_limit(self, limit: int):
        """Test hook: trip the slot-pool watchdog after `limit` supergrid cells per walk quantum (<= 0: default)."""
        self._check(self._L.uivr_debug_set_walk_limit(self._h, int(limit)), "uivr_debug_set_walk_limit")

    def set_counting(self, enable: bool):
        self._check(self._L.uivr_set_counting(self._h, int(bool(enable))), "uivr_set_counting")

    # -- device-pointer entry points (ptr = int device address, stream = int cudaStream_t) --
    @staticmethod
    def _shard(shard):
        if shard is None:
            return None
        s = Shard()
        s.rank, s.count, s.block = (int(v) for v in shard)
        return C.byref(s)

    def update_medium(self, sigma_t_ptr: int, stream: int = 0):
        self._check(self._L.uivr_update_medium(self._h, sigma_t_ptr, stream), "uivr_update_medium")

    def render_forward(self, albedo_ptr, seed, spp, image_ptr, sample_ptr=None, shard=None, stream=0):
        self._check(self._L.uivr_render_forward(self._h, albedo_ptr, seed & 0xFFFFFFFF, int(spp),
                                                self._shard(shard), image_ptr, sample_ptr, stream),
                    "uivr_render_forward")

    def render_backward(self, albedo_ptr, grad_image_ptr, seed_grad, spp_grad, dsigma_ptr, dalbedo_ptr,
                        sample_ptr=None, shard=None, stream=0):
        self._check(self._L.uivr_render_backward(self._h, albedo_ptr, grad_image_ptr,
                                                 seed_grad & 0xFFFFFFFF, int(spp_grad), self._shard(shard),
                                                 dsigma_ptr, dalbedo_ptr, sample_ptr, stream),
                    "uivr_render_backward")

    def render_forward_host(self, sigma_ptr, albedo_ptr, seed, spp, image_ptr, shard=None, stream=0):
        self._check(self._L.uivr_render_forward_host(self._h, sigma_ptr, albedo_ptr, seed & 0xFFFFFFFF,
                                                     int(spp), self._shard(shard), image_ptr, stream),
                    "uivr_render_forward_host")

    def render_backward_host(self, sigma_ptr, albedo_ptr, grad_image_ptr, seed_grad, spp_grad,
                             dsigma_ptr, dalbedo_ptr, shard=None, stream=0):
        self._check(self._L.uivr_render_backward_host(self._h, sigma_ptr, albedo_ptr, grad_image_ptr,
                                                      seed_grad & 0xFFFFFFFF, int(spp_grad),
                                                      self._shard(shard), dsigma_ptr, dalbedo_ptr, stream),
                    "uivr_render_backward_host")

    @staticmethod
    def _nerf(props: dict):
        p = NerfProps()
        p.queries_per_ray = int(props.get("queries_per_ray", 128))
        p.jittering_enabled = int(bool(props.get("jittering_enabled", True)))
        p.activation = {"identity": 0, "relu": 1}[str(props.get("activation", "identity")).lower()]
        p.hide_emitters = int(bool(props.get("hide_emitters", False)))
        return C.byref(p)

    def nerf_forward(self, props, emission_ptr, seed, spp, image_ptr, sample_ptr=None, shard=None, stream=0):
        self._check(self._L.uivr_nerf_forward(self._h, self._nerf(props), emission_ptr, seed & 0xFFFFFFFF, int(spp),
                                              self._shard(shard), image_ptr, sample_ptr, stream), "uivr_nerf_forward")

    def nerf_backward(self, props, emission_ptr, grad_image_ptr, seed_grad, spp_grad, dsigma_ptr, demission_ptr,
                      sample_ptr=None, shard=None, stream=0):
        self._check(self._L.uivr_nerf_backward(self._h, self._nerf(props), emission_ptr, grad_image_ptr,
                                               seed_grad & 0xFFFFFFFF, int(spp_grad), self._shard(shard), dsigma_ptr,
                                               demission_ptr, sample_ptr, stream), "uivr_nerf_backward")

    def adam_step(self, param_ptr, grad_ptr, m_ptr, v_ptr, n, lr, beta1, beta2, eps, t, lo, hi, stream=0):
        self._check(self._L.uivr_adam_step(self._h, param_ptr, grad_ptr, m_ptr, v_ptr, int(n), float(lr), float(beta1),
                                           float(beta2), float(eps), int(t), float(lo), float(hi), stream),
                    "uivr_adam_step")

    def upsample2x(self, in_ptr, res_xyz, channels, out_ptr, stream=0):
        r = (C.c_int32 * 3)(*[int(v) for v in res_xyz])
        self._check(self._L.uivr_upsample2x(self._h, in_ptr, r, int(channels), out_ptr, stream), "uivr_upsample2x")

    # -- instrumentation --
    def reset_counters(self, stream: int = 0):
        self._check(self._L.uivr_reset_counters(self._h, stream), "uivr_reset_counters")

    def get_counters(self, stream: int = 0) -> dict:
        out = (C.c_uint64 * len(COUNTER_NAMES))()
        self._check(self._L.uivr_get_counters(self._h, out, stream), "uivr_get_counters")
        self.check_watchdog(stream)
        return dict(zip(COUNTER_NAMES, (int(v) for v in out)))

    def kernel_ms(self, which: int) -> float:
        """CUDA-event time of the last path megakernel (0 = forward, 1 = backward)."""
        out = C.c_float()
        self._check(self._L.uivr_get_kernel_ms(self._h, int(which), C.byref(out)), "uivr_get_kernel_ms")
        return float(out.value)

    def check_watchdog(self, stream: int = 0):
        """Synchronise and raise NativeError if a persistent kernel aborted on its progress watchdog."""
        self._check(self._L.uivr_check_watchdog(self._h, None, stream), "uivr_check_watchdog")

    def launch_count(self) -> int:
        out = C.c_uint64()
        self._check(self._L.uivr_get_launch_count(self._h, C.byref(out)), "uivr_get_launch_count")
        return int(out.value)

    # -- primitive tests --
    def test_neg_log1m(self, u_ptr, n, out_ptr, stream=0):
        self._check(self._L.uivr_test_neg_log1m(self._h, u_ptr, n, out_ptr, stream), "uivr_test_neg_log1m")

    def test_atan2_turns(self, y_ptr, x_ptr, n, out_ptr, stream=0):
        self._check(self._L.uivr_test_atan2_turns(self._h, y_ptr, x_ptr, n, out_ptr, stream), "uivr_test_atan2_turns")

    def test_exp(self, x_ptr, n, out_ptr, stream=0):
        self._check(self._L.uivr_test_exp(self._h, x_ptr, n, out_ptr, stream), "uivr_test_exp")

    def test_sincos2pi(self, x_ptr, n, s_ptr, c_ptr, stream=0):
        self._check(self._L.uivr_test_sincos2pi(self._h, x_ptr, n, s_ptr, c_ptr, stream), "uivr_test_sincos2pi")

    def test_sampler(self, seed, idx0, nstreams, ndraws, out_ptr, stream=0):
        self._check(self._L.uivr_test_sampler(self._h, seed & 0xFFFFFFFF, idx0, nstreams, ndraws, out_ptr, stream),
                    "uivr_test_sampler")

    def test_sigma_lookup(self, p_ptr, n, out_ptr, stream=0):
        self._check(self._L.uivr_test_sigma_lookup(self._h, p_ptr, n, out_ptr, stream), "uivr_test_sigma_lookup")

    def get_majorant(self, out_ptr: Optional[int] = None, stream=0):
        mres = (C.c_int32 * 3)()
        self._check(self._L.uivr_get_majorant(self._h, mres, out_ptr, stream), "uivr_get_majorant")
        return tuple(int(v) for v in mres)

    def get_walk_table(self, out_ptr: Optional[int] = None, stream=0):
        """Padded supergrid + exit masks ((mres + 2)^3 uint32 words, include/uivr.h)."""
        mres = (C.c_int32 * 3)()
        self._check(self._L.uivr_get_walk_table(self._h, mres, out_ptr, stream), "uivr_get_walk_table")
        return tuple(int(v) for v in mres)
