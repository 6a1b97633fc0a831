"""Shared test inputs: seeded scenes / grids at sizes the CPU oracle finishes in seconds."""
import numpy as np

FLAG_COMBOS = {
    # opt_config.py:123-160 registry + the combo of tests/test_integrators.py:272-275
    "volpathsimple-drt": dict(use_drt=True, use_drt_subsampling=True, use_drt_mis=True),
    "volpathsimple-drt-quadratic": dict(use_drt=True, use_drt_subsampling=False, use_drt_mis=True),
    "volpathsimple-basic": dict(use_drt=False),
    "test04-nomis": dict(use_drt=True, use_drt_subsampling=True, use_drt_mis=False),
}


def hetero_grids(n, seed=7, channels_vary=True):
    """Small heterogeneous grids with empty space (numpy only, deterministic)."""
    rng = np.random.default_rng(seed)
    ax = (np.arange(n) + 0.5) / n - 0.5
    r2 = ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2
    blob = np.clip(1.0 - r2 / 0.2, 0.0, 1.0)
    noise = rng.random((n, n, n))
    sig = (blob * (0.3 + 0.7 * noise)).astype(np.float32)
    sig[sig < 0.08] = 0.0
    alb = (0.2 + 0.75 * rng.random((n, n, n, 3))).astype(np.float32)
    if not channels_vary:
        alb[..., 1] = alb[..., 0]
        alb[..., 2] = alb[..., 0]
    return sig[..., None].copy(), alb


def loss_grad(image):
    """d/d image of mean((image - 0.5)^2)  (tests/test_integrators.py:119-120)."""
    return (2.0 * (image.astype(np.float64) - 0.5) / image.size).astype(np.float32)


def rel_linf(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def random_case(uivr, seed):
    """A seeded random scene / integrator / launch configuration for the randomized parity sweep:
    anisotropic grid and box, non-square film, random camera, supergrid factor, flags, emitter."""
    import importlib
    S = importlib.import_module(uivr.__name__ + ".scene")
    rng = np.random.default_rng(1000 + seed)
    x, y, z = (int(v) for v in rng.integers(3, 34, size=3))
    az, ay, ax = [(np.arange(m) + 0.5) / m - 0.5 for m in (z, y, x)]
    r2 = az[:, None, None] ** 2 + ay[None, :, None] ** 2 + ax[None, None, :] ** 2
    sig = (np.clip(1.0 - r2 / rng.uniform(0.1, 0.3), 0.0, 1.0) * (0.2 + 0.8 * rng.random((z, y, x)))).astype(np.float32)
    sig[sig < rng.uniform(0.0, 0.15)] = 0.0
    alb = (rng.uniform(0.0, 0.3) + 0.7 * rng.random((z, y, x, 3))).astype(np.float32)
    ext = rng.uniform(0.8, 2.5, size=3)
    bmin = rng.uniform(-1.0, 0.0, size=3)
    centre = bmin + 0.5 * ext
    d = rng.standard_normal(3)
    d /= np.linalg.norm(d)
    origin = centre + d * rng.uniform(3.0, 6.0)
    target = centre + rng.uniform(-0.2, 0.2, size=3) * ext
    up = (0.0, 1.0, 0.0) if abs(d[1]) < 0.9 else (1.0, 0.0, 0.0)
    sensor = uivr.Sensor(origin=tuple(origin), target=tuple(target), up=up, fov=float(rng.uniform(15.0, 50.0)),
                         width=int(rng.integers(5, 40)), height=int(rng.integers(5, 40)))
    vol = uivr.VolumeScene(res=(x, y, z), sensor=sensor, bbox_min=tuple(bmin), bbox_extent=tuple(ext),
                           scale=float(rng.uniform(1.0, 12.0)),
                           majorant_resolution_factor=int(rng.choice([0, 2, 3, 4, 8])),
                           radiance=tuple(rng.uniform(0.1, 1.5, size=3)))
    if rng.random() < 0.4:
        img = (rng.random((int(rng.integers(2, 12)), int(rng.integers(1, 20)), 3)) ** 2).astype(np.float32)
        th = rng.uniform(0, 6.28)
        vol.envmap = S.EnvMap(img + 1e-3, scale=float(rng.uniform(0.3, 2.0)),
                              to_world=((np.cos(th), 0, np.sin(th)), (0, 1, 0), (-np.sin(th), 0, np.cos(th))))
    combo = sorted(FLAG_COMBOS)[int(rng.integers(0, len(FLAG_COMBOS)))]
    props = dict(max_depth=int(rng.integers(0, 14)), use_nee=bool(rng.random() < 0.8),
                 hide_emitters=bool(rng.random() < 0.2), **FLAG_COMBOS[combo])
    if not props.get("use_drt_subsampling", True):
        props["max_depth"] = min(props["max_depth"], 5)
    return dict(vol=vol, sig=sig[..., None].copy(), alb=alb, props=props, spp=int(rng.integers(1, 10)),
                seed=int(rng.integers(0, 2 ** 32)), seed_grad=int(rng.integers(0, 2 ** 32)),
                variant=int(rng.choice([1, 3])))
