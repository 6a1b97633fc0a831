"""Ray-batch rendering at the reference's production shape (reproduce.py:45-59): batch 32768 pixels over
32 sensors, primal spp 16 x 64 = 1024, adjoint spp 16, 256^3 grids.  Scratch timing, not the bench contract."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uivr_b200 as u

def main(n=256, batch=32768, spp=1024, spp_grad=16, reps=3):
    dev = torch.device("cuda:0")
    sig, alb = u.synthetic_grids(n)
    params = {"m.sigma_t.data": sig.to(dev).requires_grad_(True), "m.albedo.data": alb.to(dev).requires_grad_(True)}
    vol = u.benchmark_scene(n, 512, 512, scale=8.0, majorant_resolution_factor=8)
    scene = u.Scene(vol, 0)
    integ = u.get_int_config("volpathsimple-drt").create(max_depth=64)
    sensors = u.circle_sensors(32, 512, 512)
    refs = torch.rand((32, 512, 512, 3), device=dev)
    for it in range(reps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        img, si, px = u.render_batch(batch, scene, sensors, params, integ, seed=u.tea32(2 * it, 1234),
                                     seed_grad=u.tea32(2 * it + 1, 1234), spp=spp, spp_grad=spp_grad)
        e[1].record()
        loss = (img - u.gather_ref_values(refs, si, px)).abs().mean()
        loss.backward()
        e[2].record()
        torch.cuda.synchronize()
        scene.ctx.check_watchdog()
        tf, tb = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        print(f"batch {batch} spp {spp}/{spp_grad}: primal {tf:.1f} ms ({batch * spp / tf / 1e3:.0f} Msamples/s)  "
              f"adjoint {tb:.1f} ms ({batch * spp_grad / tb / 1e3:.0f} Msamples/s)  iteration {tf + tb:.1f} ms", flush=True)

if __name__ == "__main__":
    main(**{k: int(v) for k, v in (a.split("=") for a in sys.argv[1:])})
