"""Inputs of the reference-generated golden vectors (shared by make_refshim_golden.py and the
tests that consume tests/golden/refshim_*.npz).  Pure numpy + the product's scene description."""
import numpy as np

import uivr_b200 as u
from helpers import FLAG_COMBOS, hetero_grids

# registry name (python/opt_config.py:123-160) + overrides -> the reference integrator;
# the same flags as a props dict -> oracle / CUDA path
INTEGRATORS = {
    "volpathsimple-drt": ("volpathsimple-drt", {}),
    "volpathsimple-drt-quadratic": ("volpathsimple-drt-quadratic", {}),
    "volpathsimple-basic": ("volpathsimple-basic", {}),
    "test04-nomis": ("volpathsimple-drt", dict(use_drt_mis=False)),  # tests/test_integrators.py:272-275
    "no-nee": ("volpathsimple-drt", dict(use_nee=False)),
    "hide-emitters": ("volpathsimple-drt", dict(hide_emitters=True)),
    "basic-no-nee": ("volpathsimple-basic", dict(use_nee=False)),
}


def props_of(name, max_depth):
    reg, over = INTEGRATORS[name]
    p = dict(max_depth=max_depth, use_nee=True)
    p.update(FLAG_COMBOS.get(reg, {}))
    p.update(over)
    return p


# name: grid n, film w, h, spp, density scale, supergrid factor, seed, seed_grad, [(integrator, max_depth)]
CASES = {
    "cube3": dict(n=3, w=12, h=12, spp=8, scale=2.0, factor=0, seed=7, seed_grad=8,
                  runs=[(k, 8) for k in ("volpathsimple-drt", "volpathsimple-drt-quadratic",
                                         "volpathsimple-basic", "test04-nomis")]),
    "hetero12": dict(n=12, w=16, h=12, spp=4, scale=6.0, factor=4, seed=1234, seed_grad=0x38fc4d3a,
                     runs=[(k, 16) for k in ("volpathsimple-drt", "volpathsimple-drt-quadratic",
                                             "volpathsimple-basic", "test04-nomis")]),
    "hetero16": dict(n=16, w=24, h=20, spp=6, scale=10.0, factor=4, seed=99, seed_grad=777,
                     runs=[("volpathsimple-drt", 32), ("no-nee", 32), ("hide-emitters", 32),
                           ("basic-no-nee", 32), ("volpathsimple-drt", 1), ("volpathsimple-drt", 0),
                           ("volpathsimple-drt-quadratic", 2)]),
}

BATCH = dict(n=12, scale=6.0, factor=4, film=(16, 12), n_sensors=5, batch_size=64, spp=8, spp_grad=4,
             seed=4321, integrator="volpathsimple-drt", max_depth=16)


# nerf integrator (python/integrators/nerf.py): name -> (properties, offset added to the sigma_t grid;
# a negative offset gives negative raw densities, which is what `activation` is about)
NERF = dict(n=12, w=16, h=12, spp=4, scale=6.0, seed=5, seed_grad=9,
            runs={"default32": (dict(queries_per_ray=32), 0.0),
                  "nojitter17": (dict(queries_per_ray=17, jittering_enabled=False), 0.0),
                  "relu32": (dict(queries_per_ray=32, activation="relu"), -0.1),
                  "identity-negative": (dict(queries_per_ray=32), -0.1),
                  "hide128": (dict(queries_per_ray=128, hide_emitters=True), 0.0)},
            batch=dict(props=dict(queries_per_ray=24), film=(16, 12), n_sensors=5, batch_size=64, spp=8, spp_grad=4,
                       seed=4321))


def nerf_inputs(offset=0.0):
    c = NERF
    sig, em = hetero_grids(c["n"], seed=c["n"])
    vol = u.cube_test_scene(c["w"], c["h"], density_scale=c["scale"], res=(c["n"],) * 3)
    return (sig + np.float32(offset)).astype(np.float32), em, vol


# envmap emitter: the hetero12 case lit by a small lat-long map with a "sun",
# rotated about +Y; volpathsimple flag combos + the nerf integrator
ENVMAP = dict(case="hetero12", runs=[("volpathsimple-drt", 16), ("volpathsimple-basic", 16), ("no-nee", 16),
                                     ("volpathsimple-drt-quadratic", 3)],
              nerf_props=dict(queries_per_ray=16))


def test_envmap():
    from importlib import import_module
    S = import_module(u.__name__ + ".scene")
    rng = np.random.default_rng(3)
    img = (rng.random((9, 16, 3)) ** 3 * 2.0).astype(np.float32)
    img[2, 5] = (30.0, 25.0, 10.0)
    th = 0.7
    rot = ((np.cos(th), 0.0, np.sin(th)), (0.0, 1.0, 0.0), (-np.sin(th), 0.0, np.cos(th)))
    return S.EnvMap(img, scale=1.5, to_world=rot)


def envmap_inputs():
    sig, alb, vol = case_inputs(ENVMAP["case"])
    vol.envmap = test_envmap()
    return sig, alb, vol


def case_inputs(name):
    c = CASES[name]
    if name == "cube3":
        sig, alb = u.cube_test_grids()
        vol = u.cube_test_scene(c["w"], c["h"], density_scale=c["scale"])
    else:
        sig, alb = hetero_grids(c["n"], seed=c["n"])
        vol = u.cube_test_scene(c["w"], c["h"], density_scale=c["scale"], res=(c["n"],) * 3)
        vol.majorant_resolution_factor = c["factor"]
    return sig, alb, vol


def batch_inputs():
    b = BATCH
    sig, alb = hetero_grids(b["n"], seed=b["n"])
    vol = u.cube_test_scene(b["film"][0], b["film"][1], density_scale=b["scale"], res=(b["n"],) * 3)
    vol.majorant_resolution_factor = b["factor"]
    sensors = u.circle_sensors(b["n_sensors"], b["film"][0], b["film"][1])
    return sig, alb, vol, u.batched.sensor_table(sensors)


def batch_loss_grad(image):
    return (2.0 * (np.asarray(image, dtype=np.float64) - 0.5) / image.size).astype(np.float32)
