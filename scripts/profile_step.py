"""One warm-up + `reps` fwd+bwd steps of config 3 (or a smaller one) for ncu captures.
Numbers printed by a run under ncu are never bench values."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uivr_b200 as u


def main(n=256, w=512, h=512, spp=64, variant=0, factor=8, reps=1):
    dev = torch.device("cuda:0")
    sig, alb = u.synthetic_grids(n)
    sig, alb = sig.to(dev), alb.to(dev)
    vol = u.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=factor)
    scene = u.Scene(vol, 0)
    scene.ctx.set_variant(variant)
    integ = u.get_int_config("volpathsimple-drt").create(max_depth=64)
    params = {"m.sigma_t.data": sig, "m.albedo.data": alb}
    for it in range(1 + reps):
        img = integ.render(scene, params, seed=u.tea32(2 * it, 1234), spp=spp)
        g = 2 * (img - 0.5) / img.numel()
        integ.render_backward(scene, params, g, seed=u.tea32(2 * it + 1, 1234), spp=spp)
        torch.cuda.synchronize()
        print(f"step {it}: fwd {scene.ctx.kernel_ms(0):.2f} ms, bwd {scene.ctx.kernel_ms(1):.2f} ms", flush=True)


if __name__ == "__main__":
    main(**{k: int(v) for k, v in (a.split("=") for a in sys.argv[1:])})
