"""CUDA-event timings of the BASELINE.json configurations that are not the bench metric
(bench.py measures config 3): config 2 (forward-only), config 4 (32-view optimisation step with
Adam + projection + supergrid rebuild) and config 3 lit by an envmap.  One JSON line each."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import uivr_b200 as u  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def config2(reps=5):
    dev = torch.device("cuda:0")
    n, w, h, spp = 128, 256, 256, 16
    sig, _ = u.synthetic_grids(n)
    params = {"m.sigma_t.data": sig.to(dev), "m.albedo.data": torch.full((n, n, n, 3), 0.8, device=dev)}
    scene = u.Scene(u.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=8), 0)
    integ = u.get_int_config("volpathsimple-drt").create(max_depth=64)
    seeds = iter(range(1, 1000))
    ms = timed(lambda: integ.render(scene, params, seed=next(seeds), spp=spp), reps)
    print(json.dumps({"config": "2: 128^3 heterogeneous sigma_t, 256x256x16spp, forward only", "ms": ms,
                      "msamples_per_s": w * h * spp / ms / 1e3}))


def config4(reps=2, views=32, spp=32):
    dev = torch.device("cuda:0")
    n, w, h = 256, 512, 512
    sig, alb = u.synthetic_grids(n)
    tsig, talb = u.synthetic_grids(n, seed=20220722)
    sensors = u.circle_sensors(views, w, h)
    vol = u.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=8)
    scene = u.Scene(vol, 0)
    integ = u.get_int_config("volpathsimple-drt").create(max_depth=64)
    target = {"m.sigma_t.data": tsig.to(dev), "m.albedo.data": talb.to(dev)}
    refs = [integ.render(scene, target, sensor=s, seed=100 + i, spp=spp) for i, s in enumerate(sensors)]
    params = {"m.sigma_t.data": sig.to(dev), "m.albedo.data": alb.to(dev)}
    opt = u.Adam(5e-3, params)                                   # reproduce.py:50
    opt.set_learning_rate(u.learning_rates(5e-3, list(params), 0, 100, None, {"m.albedo.data": 2.0}))  # scene_config.py:67-71
    grads = {k: torch.zeros_like(p) for k, p in params.items()}
    it = iter(range(1000))
    losses = []
    ms = timed(lambda: losses.append(u.optimization_step(scene, integ, opt, sensors, refs, next(it), spp, grads=grads)), reps)
    print(json.dumps({"config": f"4: 256^3, {views} views x 512x512x{spp}spp, render + L1 + backward per view, Adam + clamp + supergrid rebuild",
                      "ms_per_step": ms, "steps_per_s": 1e3 / ms, "msamples_per_s": views * w * h * spp / ms / 1e3,
                      "losses": losses}))
    # the same with the views alternating between two contexts / streams (the drain of one view's launches overlaps
    # the other view's kernels)
    for n_lanes in (2, 3):
        lanes = u.make_view_lanes(scene, params, n_lanes)
        losses2 = []
        ms2 = timed(lambda: losses2.append(u.optimization_step(scene, integ, opt, sensors, refs, next(it), spp, grads=grads, lanes=lanes)), reps)
        print(json.dumps({"config": f"4 with {n_lanes} view lanes: 256^3, {views} views x 512x512x{spp}spp", "ms_per_step": ms2,
                          "steps_per_s": 1e3 / ms2, "msamples_per_s": views * w * h * spp / ms2 / 1e3, "losses": losses2}))
        del lanes


def config3_envmap(reps=3):
    from importlib import import_module
    S = import_module(u.__name__ + ".scene")
    dev = torch.device("cuda:0")
    n, w, h, spp = 256, 512, 512, 64
    sig, alb = u.synthetic_grids(n)
    params = {"m.sigma_t.data": sig.to(dev), "m.albedo.data": alb.to(dev)}
    rng = np.random.default_rng(0)
    img = (rng.random((1024, 2048, 3)) ** 4).astype(np.float32)
    img[200:220, 500:520] = (500.0, 450.0, 300.0)   # a sun
    vol = u.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=8)
    vol.envmap = S.EnvMap(img, scale=1.0)
    scene = u.Scene(vol, 0)
    integ = u.get_int_config("volpathsimple-drt").create(max_depth=64)
    seeds = iter(range(1, 1000))
    state = {}

    def step():
        s = next(seeds)
        state["img"] = integ.render(scene, params, seed=s, spp=spp)
        g = 2 * (state["img"] - 0.5) / state["img"].numel()
        integ.render_backward(scene, params, g, seed=u.tea32(s, 1), spp=spp)
    ms = timed(step, reps)
    print(json.dumps({"config": "3 + envmap (2048x1024 lat-long map with a sun): 256^3, 512x512x64spp, DRT fwd+bwd",
                      "ms_per_step": ms, "msamples_per_s": w * h * spp / ms / 1e3,
                      "fwd_kernel_ms": scene.ctx.kernel_ms(0), "bwd_kernel_ms": scene.ctx.kernel_ms(1),
                      "img_mean": float(state["img"].mean())}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["2", "4", "env"]
    if "2" in which:
        config2()
    if "4" in which:
        config4()
    if "env" in which:
        config3_envmap()
