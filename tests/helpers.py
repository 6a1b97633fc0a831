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
