"""Tiny forward + backward (slot-pool and one-sample-per-lane kernels), ray batch, envmap, nerf, and a scene whose
shadow walks overflow the NEE collision log, for compute-sanitizer runs."""
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
for variant in (3, 1):
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
# nerf integrator + envmap emitter (slot-pool ENV instances, one-sample-per-lane kernels, nerf kernels)
from importlib import import_module
S = import_module(u.__name__ + ".scene")
rng = np.random.default_rng(0)
envimg = (rng.random((17, 32, 3)) ** 3).astype(np.float32)
envimg[4, 7] = (50.0, 40.0, 30.0)
vol_e = u.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=4)
vol_e.envmap = S.EnvMap(envimg, scale=0.8)
for variant in (3, 1):
    scene = u.Scene(vol_e, 0)
    scene.ctx.set_variant(variant)
    img = integ.render(scene, params, seed=1, spp=spp)
    g = 2 * (img - 0.5) / img.numel()
    ds, da = integ.render_backward(scene, params, g, seed=2, spp=spp)
    torch.cuda.synchronize()
    scene.ctx.check_watchdog()
    print("envmap variant", variant, float(img.mean()), float(ds.abs().sum()), float(da.abs().sum()))
nerf = u.get_int_config("nerf").create(max_depth=4, queries_per_ray=48)
pn = {"m.sigma_t.data": params["m.sigma_t.data"], "m.emission.data": params["m.albedo.data"]}
for v in (vol, vol_e):
    scene = u.Scene(v, 0)
    img = nerf.render(scene, pn, seed=1, spp=spp)
    ds, de = nerf.render_backward(scene, pn, 2 * (img - 0.5) / img.numel(), seed=2, spp=spp)
    torch.cuda.synchronize()
    print("nerf", "envmap" if v.envmap is not None else "constant", float(img.mean()), float(ds.abs().sum()), float(de.abs().sum()))
# thin medium under a high majorant: shadow walks with > 32 null collisions (log overflow -> second walk), odd grid sizes
rng = np.random.default_rng(5)
m = 15
sig_s = (0.01 + 0.02 * rng.random((m, m, m, 1))).astype(np.float32)
sig_s[m // 2, m // 2, m // 2, 0] = 1.0
alb_s = (0.3 + 0.6 * rng.random((m, m, m, 3))).astype(np.float32)
vol_s = u.benchmark_scene(m, 24, 20, scale=60.0, majorant_resolution_factor=16)
scene = u.Scene(vol_s, 0)
ps = {"m.sigma_t.data": torch.from_numpy(sig_s).to(dev), "m.albedo.data": torch.from_numpy(alb_s).to(dev)}
integ6 = u.get_int_config("volpathsimple-drt").create(max_depth=6)
img = integ6.render(scene, ps, seed=1, spp=4)
ds, da = integ6.render_backward(scene, ps, 2 * (img - 0.5) / img.numel(), seed=2, spp=4)
torch.cuda.synchronize()
scene.ctx.check_watchdog()
print("log overflow scene", float(img.mean()), float(ds.abs().sum()), float(da.abs().sum()))
