#!/usr/bin/env python
"""End-to-end use of the drop-in path, shaped like python/optimize.py:run_optimization:
recover a synthetic medium from rendered reference views with ray-batch rendering (batched.py),
L1 loss, Adam + projection, Last25 schedule, coarse-to-fine upsampling and `.vol` checkpoints.

    python examples/optimize_synthetic.py [n_iter=200] [res=16] [batch=8192] [out=/tmp/uivr_opt]

Prints one JSON line per logged iteration (the reference never logs its loss).
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import uivr_b200 as u  # noqa: E402


def main(n_iter=200, res=16, batch=8192, spp=16, primal_spp_factor=4, n_sensors=16, film=128, out="/tmp/uivr_opt"):
    dev = torch.device("cuda:0")
    final_res = res * 4                                   # two x2 upsamplings
    upsample_at = u.upsample_iterations([0.2, 0.5], n_iter)
    sensors = u.circle_sensors(n_sensors, film, film)
    integ = u.get_int_config("volpathsimple-drt").create(max_depth=32)

    # reference images of the target medium (render_reference_image, optimize.py:24-53)
    sig_t, alb_t = u.synthetic_grids(final_res)
    target = {"medium1.sigma_t.data": sig_t.to(dev), "medium1.albedo.data": alb_t.to(dev)}
    tscene = u.Scene(u.benchmark_scene(final_res, film, film, scale=8.0,
                                       majorant_resolution_factor=u.adjust_majorant_res_factor(8, (final_res,) * 3)), 0)
    refs = torch.stack([integ.render(tscene, target, sensor=s, seed=1000 + i, spp=256) for i, s in enumerate(sensors)])

    # initial state (initialize_scene, optimize.py:134-166): constant grids at the coarse resolution
    params = {"medium1.sigma_t.data": torch.full((res, res, res, 1), 0.1, device=dev),
              "medium1.albedo.data": torch.full((res, res, res, 3), 0.6, device=dev)}
    scene = u.Scene(u.benchmark_scene(res, film, film, scale=8.0,
                                      majorant_resolution_factor=u.adjust_majorant_res_factor(8, (res,) * 3)), 0)
    opt = u.Adam(lr=5e-3, params=params)
    lr_factors = {"medium1.albedo.data": 2.0}             # scene_config.py:67-71
    u.save_params(os.path.join(out, "params"), opt.params, "initial")

    t0 = time.perf_counter()
    for it in range(n_iter):
        seed, seed_grad = u.tea32(2 * it, 1234), u.tea32(2 * it + 1, 1234)          # optimize.py:327-328
        opt.set_learning_rate(u.learning_rates(5e-3, list(opt.params), it, n_iter, "last25", lr_factors))
        if it in upsample_at:                                                        # optimize.py:330
            shapes = u.upsample_params(scene, opt, 8)
            print(json.dumps({"it": it, "upsampled": {k: list(v) for k, v in shapes.items()}}), flush=True)
        p = {k: v.requires_grad_(True) for k, v in opt.params.items()}
        image, si, px = u.render_batch(batch, scene, sensors, p, integ, seed=seed, seed_grad=seed_grad,
                                       spp=spp * primal_spp_factor, spp_grad=spp)    # optimize.py:334-340
        loss = (image - u.gather_ref_values(refs, si, px)).abs().mean()              # losses.l1
        loss.backward()                                                              # optimize.py:350
        grads = {k: v.grad for k, v in p.items()}
        for v in p.values():
            v.requires_grad_(False)
        opt.step(scene.ctx, grads, max_density=250.0)                                # optimize.py:352-353
        scene.update_medium(opt.params["medium1.sigma_t.data"], force=True)          # optimize.py:354
        for v in opt.params.values():
            v.grad = None
        if it % 20 == 0 or it == n_iter - 1:
            print(json.dumps({"it": it, "loss": float(loss.detach()), "res": list(opt.params["medium1.sigma_t.data"].shape[:3]),
                              "elapsed_s": round(time.perf_counter() - t0, 2)}), flush=True)
    scene.ctx.check_watchdog()
    files = u.save_params(os.path.join(out, "params"), opt.params, "final")
    print(json.dumps({"done": True, "checkpoints": sorted(files.values())}), flush=True)


if __name__ == "__main__":
    kw = {}
    for a in sys.argv[1:]:
        k, v = a.split("=")
        kw[k] = v if k == "out" else int(v)
    main(**kw)
