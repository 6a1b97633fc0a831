"""Timing of the `nerf` integrator (python/integrators/nerf.py) on the config-3 shapes:
256^3 sigma_t + 256^3 x 3 emission, 512 x 512 x spp, queries_per_ray 128, forward + backward.
Prints one JSON line with the CUDA-event kernel times and the algorithmic bytes from the event
counters (32 B per sigma_t tap, 96 B per emission tap, 64 / 192 B per gradient scatter event)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import uivr_b200 as u  # noqa: E402


def main(n=256, w=512, h=512, spp=16, queries=128, reps=3):
    dev = torch.device("cuda:0")
    sig, em = u.synthetic_grids(n)
    sig, em = sig.to(dev), em.to(dev)
    vol = u.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=8)
    scene = u.Scene(vol, 0)
    integ = u.get_int_config("nerf").create(max_depth=4, queries_per_ray=queries)
    params = {"m.sigma_t.data": sig, "m.emission.data": em}
    S = w * h * spp
    tf = tb = 0.0
    for it in range(reps + 1):
        img = integ.render(scene, params, seed=1234 + it, spp=spp)
        g = 2 * (img - 0.5) / img.numel()
        ds, de = integ.render_backward(scene, params, g, seed=u.tea32(1234 + it, 1), spp=spp)
        torch.cuda.synchronize()
        if it:  # first iteration = warm-up
            tf += scene.ctx.kernel_ms(0) / reps
            tb += scene.ctx.kernel_ms(1) / reps
    scene.ctx.set_counting(True)
    scene.ctx.reset_counters()
    integ.render(scene, params, seed=1234, spp=spp)
    cf = scene.ctx.get_counters()
    scene.ctx.reset_counters()
    integ.render_backward(scene, params, g, seed=u.tea32(1234, 1), spp=spp)
    cb = scene.ctx.get_counters()
    bytes_of = lambda c: 32 * c["sigma_taps"] + 96 * c["albedo_taps"] + 64 * c["sigma_scatters"] + 192 * c["albedo_scatters"]
    peak = 6546.2
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {"workload": f"nerf {n}^3, {w}x{h}x{spp}spp, {queries} queries/ray", "samples": S,
           "fwd_ms": tf, "bwd_ms": tb, "msamples_per_s": S / (tf + tb) / 1e3,
           "fwd_alg_bytes": bytes_of(cf), "bwd_alg_bytes": bytes_of(cb),
           "fwd_alg_gbs": bytes_of(cf) / tf / 1e6, "bwd_alg_gbs": bytes_of(cb) / tb / 1e6,
           "hbm_peak_gbs": peak, "fwd_frac": bytes_of(cf) / tf / 1e6 / peak, "bwd_frac": bytes_of(cb) / tb / 1e6 / peak,
           "camera_hit_fraction": cf["camera_hits"] / S, "img_mean": float(img.mean()),
           "abs_dsigma": float(ds.abs().sum()), "abs_demission": float(de.abs().sum())}
    print(json.dumps(out))


if __name__ == "__main__":
    kw = {}
    for a in sys.argv[1:]:
        k, v = a.split("=")
        kw[k] = int(v)
    main(**kw)
