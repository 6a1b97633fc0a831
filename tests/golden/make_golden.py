"""Generate the golden fixtures of tests/golden/ with the CPU oracle.

The reference ships no golden vectors for this path (tests/test_integrators.py
regenerates everything and its gradient assertions are disabled) and cannot be imported here
(Mitsuba 3 / Dr.Jit absent), so these fixtures pin OUR restatement: they freeze the oracle's
outputs (per-sample radiance bit patterns, image, gradients, event counters) on seeded inputs so
that (a) the oracle cannot drift silently and (b) the CUDA path is compared against committed
numbers, not only against a freshly built checker.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import uivr_b200 as u  # noqa: E402
from helpers import FLAG_COMBOS, hetero_grids, loss_grad  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {
    # name: (grid n, film w, h, spp, density scale, supergrid factor, max_depth, seed, seed_grad)
    "hetero12": (12, 16, 12, 4, 6.0, 4, 16, 1234, 0x38fc4d3a),
    "cube3": (3, 12, 12, 8, 2.0, 0, 8, 7, 8),
}


def case_inputs(name):
    n, w, h, spp, scale, factor, max_depth, seed, seed_grad = CASES[name]
    if name == "cube3":
        sig, alb = u.cube_test_grids()
        vol = u.cube_test_scene(w, h, density_scale=scale)
    else:
        sig, alb = hetero_grids(n, seed=n)
        vol = u.cube_test_scene(w, h, density_scale=scale, res=(n, n, n))
        vol.majorant_resolution_factor = factor
    return sig, alb, vol, spp, max_depth, seed, seed_grad


def main():
    O.build()
    for name in CASES:
        sig, alb, vol, spp, max_depth, seed, seed_grad = case_inputs(name)
        desc = vol.as_dict()
        out = {"sigma_t": sig, "albedo": alb}
        for combo, flags in FLAG_COMBOS.items():
            props = dict(max_depth=max_depth, use_nee=True, **flags)
            img, samples, cf = O.render_forward(desc, props, sig, alb, seed, spp, want_samples=True, nthreads=1)
            gimg = loss_grad(img)
            ds, da, samples_g, cb = O.render_backward(desc, props, sig, alb, gimg, seed_grad, spp,
                                                      want_samples=True, nthreads=1)
            out[f"{combo}/image"] = img
            out[f"{combo}/samples"] = samples.view(np.uint32)
            out[f"{combo}/samples_grad_pass"] = samples_g.view(np.uint32)
            out[f"{combo}/dsigma"] = ds
            out[f"{combo}/dalbedo"] = da
            out[f"{combo}/counters_fwd"] = np.array([cf[k] for k in O.COUNTER_NAMES], dtype=np.uint64)
            out[f"{combo}/counters_bwd"] = np.array([cb[k] for k in O.COUNTER_NAMES], dtype=np.uint64)
            cp = O.last_backward_primal_counters()   # the share of the primal pass inside the backward
            out[f"{combo}/counters_bwd_primal"] = np.array([cp[k] for k in O.COUNTER_NAMES], dtype=np.uint64)
            cr = O.last_backward_replay_counters()   # ... and of the second NEE walks
            out[f"{combo}/counters_bwd_replay"] = np.array([cr[k] for k in O.COUNTER_NAMES], dtype=np.uint64)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, "written")
    # integer known answers (KA5)
    ka = {
        "pcg32_42_54": O.pcg32_stream(42, 54, 6),
        "pcg32_default": O.pcg32_stream(0x853c49e6748fea9b, 0xda3e39cb94b95bdb, 3),
        "tea_0_0": np.array(O.tea(0, 0), dtype=np.uint32),
        "tea_1234_1": np.array(O.tea(1234, 1), dtype=np.uint32),
        "sampler_1234_0": O.sampler_floats(1234, 0, 8).view(np.uint32),
        "alt_seed_0x38fc4d3a": np.array([O.alt_seed(0x38fc4d3a)], dtype=np.uint32),
    }
    np.savez_compressed(os.path.join(HERE, "rng.npz"), **ka)
    print("rng written")


if __name__ == "__main__":
    main()
