#!/usr/bin/env python
"""The reference's whole optimisation run (python/optimize.py:run_optimization, configured as
python/reproduce.py:44-59 does) on a synthetic medium, through the drop-in surface:

    python examples/run_optimization_synthetic.py [res=64] [n_iter=400] [batch=32768] [film=128] [out=outputs/synthetic]

Reference views are rendered once into <out>/references/ref_%06d.exr and reused; checkpoints go to
<out>/params/*.vol, previews to <out>/opt_*.exr.  One JSON line per logged iteration.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uivr_b200 as u  # noqa: E402

KEYS = ["medium1.sigma_t.data", "medium1.albedo.data"]


def main(res=64, n_iter=400, batch=32768, film=128, n_sensors=16, out="outputs/synthetic"):
    sig_ref, alb_ref = u.synthetic_grids(res)                      # the medium to recover (SURVEY §8d recipe)
    scene_config = u.SceneConfig(
        "synthetic", volume=u.benchmark_scene(res, film, film, scale=8.0, majorant_resolution_factor=8),
        scene_sensors=u.circle_sensors(n_sensors, film, film), param_keys=KEYS, sensors=list(range(n_sensors)),
        start_from_value={KEYS[0]: 0.04, KEYS[1]: 0.6}, ref_params={KEYS[0]: sig_ref, KEYS[1]: alb_ref},
        ref_spp=1024, max_depth=64, references=os.path.join(out, "references"))
    os.makedirs(scene_config.references, exist_ok=True)
    opt_config = u.OptimizationConfig(
        "synthetic-drt", spp=16, n_iter=n_iter, lr=5e-3, primal_spp_factor=64, batch_size=batch or None,
        lr_schedule=u.Schedule.Last25, upsample=[0.04, 0.16, 0.36] if res >= 64 else None,
        render_initial=False, preview_stride=max(1, n_iter // 4), preview_spp=256, checkpoint_stride=None)
    t0 = time.perf_counter()

    def log(it, loss):
        if it % max(1, n_iter // 20) == 0 or it == n_iter - 1:
            print(json.dumps({"it": it, "loss": round(loss, 6), "elapsed_s": round(time.perf_counter() - t0, 2)}), flush=True)

    scene, params, opt = u.run_optimization(out, opt_config, scene_config, "volpathsimple-drt", callback=log)
    print(json.dumps({"done": True, "resolution": list(params[KEYS[0]].shape[:3]),
                      "checkpoints": sorted(os.listdir(os.path.join(out, "params")))}), flush=True)


if __name__ == "__main__":
    kw = {}
    for a in sys.argv[1:]:
        k, v = a.split("=")
        kw[k] = v if k == "out" else int(v)
    main(**kw)
