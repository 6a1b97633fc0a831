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


def run_host():
    """Host-side functions of python/optimize.py, python/opt_config.py and python/losses.py, run unmodified."""
    import types
    ref = R.load_reference()
    oc, opt_mod = ref.opt_config, ref.optimize
    keys = ["m.sigma_t.data", "m.albedo.data", "m.emission.data"]
    sc = types.SimpleNamespace(param_lr_factors={"m.albedo.data": 2.0}, param_keys=keys, max_density=250.0,
                               majorant_resolution_factor=8)
    out = {}
    for name, sched in (("last25", oc.Schedule.Last25), ("constant", oc.Schedule.Constant), ("none", None)):
        cfg = oc.OptimizationConfig("x", spp=4, n_iter=101, lr=5e-3, lr_schedule=sched)
        out[f"lr/{name}"] = np.array([[cfg.learning_rates(sc, it)[k] for k in keys] for it in range(101)])
    for i, (ups, n_iter) in enumerate((([0.25, 0.5, 0.75], 101), ([0.0, 1.0, 0.333], 3000), (None, 10))):
        cfg = oc.OptimizationConfig("x", spp=4, n_iter=n_iter, lr=1.0, upsample=ups)
        out[f"upsample_at/{i}"] = np.array(sorted(cfg.upsample_at), dtype=np.int64)
        out[f"should_upsample/{i}"] = np.array([cfg.should_upsample(it) for it in range(n_iter)])
    rng = np.random.default_rng(0)
    vals = {k: (rng.standard_normal(200) * (300.0 if "sigma" in k else 2.0)).astype(np.float32) for k in keys}
    opt = {k: R.Tensor(v) for k, v in vals.items()}
    opt_mod.enforce_valid_params(sc, opt)
    for k in keys:
        out[f"clip_in/{k}"], out[f"clip_out/{k}"] = vals[k], opt[k].v
    table = []
    for factor in (0, 1, 2, 3, 4, 8, 16):
        for res in ((3, 3, 3), (8, 8, 8), (15, 16, 17), (16, 16, 16), (31, 40, 64), (64, 64, 64), (256, 256, 256), (8, 64, 64)):
            sc.majorant_resolution_factor = factor
            R.Medium._majorant_resolution_factor = -1
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                opt_mod.adjust_majorant_res_factor(sc, R.Scene(), (*res, 1))
            table.append((factor, *res, R.Medium._majorant_resolution_factor))
    out["majorant_factor_table"] = np.array(table, dtype=np.int64)
    for i, shp in enumerate(((3, 3, 3, 1), (2, 5, 4, 3), (1, 4, 2, 1))):
        g = rng.random(shp).astype(np.float32)
        new = (*[2 * r for r in shp[:3]], shp[3])
        out[f"upsample_in/{i}"], out[f"upsample_out/{i}"] = g, opt_mod.upsample_grid(R.Tensor(g), shp, new).v
    refs = rng.random((4, 6, 7, 3)).astype(np.float32)
    si, px, py = rng.integers(0, 4, 50), rng.integers(0, 7, 50), rng.integers(0, 6, 50)
    got = opt_mod.gather_ref_values(R.Tensor(refs), R.UInt32(si), R.VecU(R.UInt32(px), R.UInt32(py)))
    out.update({"gather/refs": refs, "gather/sensor_idx": si, "gather/pixels": np.stack([px, py], axis=1), "gather/out": got.v})
    a, b = rng.random((5, 6, 3)).astype(np.float32), rng.random((5, 6, 3)).astype(np.float32)
    out.update({"loss/a": a, "loss/b": b, "loss/l1": ref.losses.l1(R.Tensor(a), R.Tensor(b)).v})
    return out


LOSS_NAMES = ("average", "l1", "l2", "root_mean_squared_error", "huber", "mean_relative_absolute_error",
              "mean_relative_squared_error", "root_mean_relative_squared_error", "psnr")


def run_losses():
    """Every function of python/losses.py, run unmodified, on two image pairs (one with residuals beyond the
    huber delta on both sides); written to refshim_losses.npz."""
    ref = R.load_reference()
    rng = np.random.default_rng(7)
    out = {}
    for tag, scale in (("unit", 1.0), ("wide", 4.0)):
        a = (rng.random((5, 6, 3)) * scale).astype(np.float32)
        b = (rng.random((5, 6, 3)) * scale).astype(np.float32)
        out[f"{tag}/a"], out[f"{tag}/b"] = a, b
        for name in LOSS_NAMES:
            out[f"{tag}/{name}"] = np.asarray(getattr(ref.losses, name)(R.Tensor(a), R.Tensor(b)).v, dtype=np.float32).reshape(-1)
        out[f"{tag}/huber_delta_half"] = ref.losses.huber(R.Tensor(a), R.Tensor(b), delta=0.5).v.reshape(-1)
        out[f"{tag}/psnr_max4"] = ref.losses.psnr(R.Tensor(a), R.Tensor(b), max_value=4.0).v.reshape(-1)
        out[f"{tag}/mrse_eps"] = ref.losses.mean_relative_squared_error(R.Tensor(a), R.Tensor(b), epsilon=0.1).v.reshape(-1)
    return out


RANDOM_CASES = 16  # the seeds of tests/test_gpu_parity.py::test_randomized_parity_sweep


def run_random():
    """The randomized sweep of helpers.random_case (anisotropic grids / boxes, random cameras and films,
    supergrid factors, every flag, constant and envmap emitters) through the reference's files."""
    from helpers import random_case
    import uivr_b200 as u
    out = {}
    for case in range(RANDOM_CASES):
        c = random_case(u, case)
        props = dict(c["props"])
        integ = R.make_integrator("volpathsimple-drt", max_depth=props.pop("max_depth"), **props)
        desc = c["vol"].as_dict()
        img, samples = R.render_forward(desc, integ, c["sig"], c["alb"], c["seed"], c["spp"])
        gimg = loss_grad(img)
        ds, da, samples_g = R.render_backward(desc, integ, c["sig"], c["alb"], gimg, c["seed_grad"], c["spp"])
        out.update({f"{case}/image": img, f"{case}/samples": samples, f"{case}/grad_image": gimg,
                    f"{case}/samples_grad_pass": samples_g, f"{case}/dsigma": ds.astype(np.float32),
                    f"{case}/dalbedo": da.astype(np.float32)})
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
    np.savez_compressed(os.path.join(HERE, "refshim_host.npz"), **run_host())
    print("host written")
    np.savez_compressed(os.path.join(HERE, "refshim_losses.npz"), **run_losses())
    print("losses written")
    np.savez_compressed(os.path.join(HERE, "refshim_random.npz"), **run_random())
    print("random sweep written")
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
