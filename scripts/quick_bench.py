"""Scratch timing of fwd+bwd on a BASELINE config (not the bench contract; see bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uivr_b200 as u

def main(n=256, w=512, h=512, spp=64, variant=1, factor=8, reps=3, counters=1, depth=64):
    dev = torch.device("cuda:0")
    sig, alb = u.synthetic_grids(n)
    sig, alb = sig.to(dev), alb.to(dev)
    vol = u.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=factor)
    scene = u.Scene(vol, 0)
    scene.ctx.set_variant(variant)
    integ = u.get_int_config("volpathsimple-drt").create(max_depth=depth)
    params = {"m.sigma_t.data": sig, "m.albedo.data": alb}
    S = w * h * spp
    for it in range(reps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        img = integ.render(scene, params, seed=1234, spp=spp)
        e[1].record()
        g = 2 * (img - 0.5) / img.numel()
        ds, da = integ.render_backward(scene, params, g, seed=u.tea32(1234, 1), spp=spp)
        e[2].record()
        torch.cuda.synchronize()
        tf, tb = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        print(f"n={n} {w}x{h}x{spp} depth={depth} variant={variant} factor={factor}: fwd {tf:.1f} ms  bwd {tb:.1f} ms  "
              f"-> {S / (tf + tb) / 1e3:.1f} Msamples/s  img mean {img.mean().item():.4f} "
              f"|ds| {ds.abs().sum().item():.4e} |da| {da.abs().sum().item():.4e}", flush=True)
    if not counters:
        return
    scene.ctx.set_counting(True); scene.ctx.reset_counters()
    img = integ.render(scene, params, seed=1234, spp=spp)
    cf = scene.ctx.get_counters()
    scene.ctx.reset_counters()
    integ.render_backward(scene, params, g, seed=u.tea32(1234, 1), spp=spp)
    cb = scene.ctx.get_counters()
    print("fwd counters/sample", {k: round(v / S, 2) for k, v in cf.items()})
    print("bwd counters/sample", {k: round(v / S, 2) for k, v in cb.items()})

if __name__ == "__main__":
    kw = {}
    for a in sys.argv[1:]:
        k, v = a.split("=")
        kw[k] = int(v)
    main(**kw)
