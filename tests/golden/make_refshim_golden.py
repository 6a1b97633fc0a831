"""Generate tests/golden/refshim_*.npz by running the REFERENCE's own python files.

    python tests/golden/make_refshim_golden.py      # needs /root/reference (this container only)

oracle/refshim.py imports /root/reference/python/integrators/volpathsimple.py, batched.py and
opt_config.py UNMODIFIED on top of a numpy stand-in for Mitsuba 3 / Dr.Jit (which cannot be
installed here).  The vectors written below are therefore outputs of the reference's control flow,
masks, RNG draw order and gradient formulae; the upstream (un-vendored Mitsuba branch) arithmetic
under them is the oracle's (see the header of oracle/refshim.py for exactly what is pinned).
The oracle and the CUDA path are then tested against these files (tests/test_refshim_golden.py,
tests/test_gpu_parity.py) -- the reference cannot travel to the GPU box, the vectors can.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from helpers import loss_grad  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import refshim as R  # noqa: E402
import refshim_cases as RC  # noqa: E402


def run_case(name):
    c = RC.CASES[name]
    sig, alb, vol = RC.case_inputs(name)
    desc = vol.as_dict()
    out = {"sigma_t": sig, "albedo": alb}
    for integ_name, max_depth in c["runs"]:
        reg, over = RC.INTEGRATORS[integ_name]
        integ = R.make_integrator(reg, max_depth=max_depth, **over)
        key = f"{integ_name}@{max_depth}"
        img, samples = R.render_forward(desc, integ, sig, alb, c["seed"], c["spp"])
        gimg = loss_grad(img)
        ds, da, samples_g = R.render_backward(desc, integ, sig, alb, gimg, c["seed_grad"], c["spp"])
        out[f"{key}/image"] = img
        out[f"{key}/samples"] = samples
        out[f"{key}/grad_image"] = gimg
        out[f"{key}/samples_grad_pass"] = samples_g
        out[f"{key}/dsigma"] = ds
        out[f"{key}/dalbedo"] = da
    return out


def run_batch():
    b = RC.BATCH
    sig, alb, vol, tab = RC.batch_inputs()
    reg, over = RC.INTEGRATORS[b["integrator"]]
    integ = R.make_integrator(reg, max_depth=b["max_depth"], **over)
    res = R.render_batch(vol.as_dict(), integ, sig, alb, tab, b["film"], b["batch_size"], b["seed"], b["spp"],
                         spp_grad=b["spp_grad"], grad_image_fn=RC.batch_loss_grad)
    res["sensors"] = tab
    return res


def run_envmap():
    e = RC.ENVMAP
    c = RC.CASES[e["case"]]
    sig, alb, vol = RC.envmap_inputs()
    desc = vol.as_dict()
    out = {}
    for integ_name, max_depth in e["runs"]:
        reg, over = RC.INTEGRATORS[integ_name]
        integ = R.make_integrator(reg, max_depth=max_depth, **over)
        key = f"{integ_name}@{max_depth}"
        img, samples = R.render_forward(desc, integ, sig, alb, c["seed"], c["spp"])
        gimg = loss_grad(img)
        ds, da, samples_g = R.render_backward(desc, integ, sig, alb, gimg, c["seed_grad"], c["spp"])
        out.update({f"{key}/image": img, f"{key}/samples": samples, f"{key}/grad_image": gimg,
                    f"{key}/samples_grad_pass": samples_g, f"{key}/dsigma": ds, f"{key}/dalbedo": da})
    integ = R.make_integrator("nerf", max_depth=4, **e["nerf_props"])
    img, samples = R.render_forward(desc, integ, sig, alb, c["seed"], c["spp"])
    out.update({"nerf/image": img, "nerf/samples": samples})
    return out


def run_nerf():
    c = RC.NERF
    out = {}
    for name, (props, offset) in c["runs"].items():
        sig, em, vol = RC.nerf_inputs(offset)
        desc = vol.as_dict()
        integ = R.make_integrator("nerf", max_depth=4, **props)  # opt_config.py:162-169 registry entry
        img, samples = R.render_forward(desc, integ, sig, em, c["seed"], c["spp"])
        gimg = loss_grad(img)
        ds, de, samples_g = R.render_backward(desc, integ, sig, em, gimg, c["seed_grad"], c["spp"])
        out.update({f"{name}/image": img, f"{name}/samples": samples, f"{name}/grad_image": gimg,
                    f"{name}/samples_grad_pass": samples_g, f"{name}/dsigma": ds, f"{name}/demission": de})
    # the same integrator under python/batched.py render_batch
    b = c["batch"]
    sig, em, vol = RC.nerf_inputs()
    tab = RC.batch_inputs()[3]
    integ = R.make_integrator("nerf", max_depth=4, **b["props"])
    res = R.render_batch(vol.as_dict(), integ, sig, em, tab, b["film"], b["batch_size"], b["seed"], b["spp"],
                         spp_grad=b["spp_grad"], grad_image_fn=RC.batch_loss_grad)
    out.update({"batch/image": res["image"], "batch/dsigma": res["dsigma"], "batch/demission": res["dalbedo"],
                "batch/sensors": tab})
    return out


def main():
    O.build()
    np.savez_compressed(os.path.join(HERE, "refshim_envmap.npz"), **run_envmap())
    print("envmap written")
    np.savez_compressed(os.path.join(HERE, "refshim_nerf.npz"), **run_nerf())
    print("nerf written")
    for name in RC.CASES:
        np.savez_compressed(os.path.join(HERE, f"refshim_{name}.npz"), **run_case(name))
        print(name, "written")
    np.savez_compressed(os.path.join(HERE, "refshim_batch.npz"), **run_batch())
    print("batch written")


if __name__ == "__main__":
    main()
