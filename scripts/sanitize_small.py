"""Tiny forward + split backward + combined backward + ray batch, for compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import uivr_b200 as u

dev = torch.device("cuda:0")
n, w, h, spp = 16, 48, 48, 8
sig, alb = u.synthetic_grids(n)
params = {"m.sigma_t.data": sig.to(dev), "m.albedo.data": alb.to(dev)}
vol = u.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=4)
integ = u.get_int_config("volpathsimple-drt").create(max_depth=16)
for variant in (3, 2):
    scene = u.Scene(vol, 0)
    scene.ctx.set_variant(variant)
    img = integ.render(scene, params, seed=1, spp=spp)
    g = 2 * (img - 0.5) / img.numel()
    ds, da = integ.render_backward(scene, params, g, seed=2, spp=spp)
    torch.cuda.synchronize()
    scene.ctx.check_watchdog()
    print("variant", variant, float(img.mean()), float(ds.abs().sum()), float(da.abs().sum()))
scene = u.Scene(vol, 0)
p2 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
image, si, px = u.render_batch(512, scene, u.circle_sensors(3, 32, 32), p2, integ, seed=9, spp=4)
image.mean().backward()
torch.cuda.synchronize()
scene.ctx.check_watchdog()
print("batch", float(image.mean()), float(p2["m.sigma_t.data"].grad.abs().sum()))
