"""Kernel time of ONE rank's share of an N-way pixel-sharded config-3 step on a single GPU
(no communication): isolates what the shard block size does to the kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uivr_b200 as u

def main(count=4, spp=256, reps=3):
    dev = torch.device("cuda:0")
    sig, alb = u.synthetic_grids(256)
    sig, alb = sig.to(dev), alb.to(dev)
    vol = u.benchmark_scene(256, 512, 512, scale=8.0, majorant_resolution_factor=8)
    integ = u.get_int_config("volpathsimple-drt").create(max_depth=64)
    params = {"m.sigma_t.data": sig, "m.albedo.data": alb}
    for block in (1, 16, 64, 512, 4096, 65536):
        for rank in (0, count - 1):
            scene = u.Scene(vol, 0)
            shard = (rank, count, block) if count > 1 else None
            for it in range(reps):
                img = integ.render(scene, params, seed=10 + it, spp=spp, shard=shard)
                g = 2 * (img - 0.5) / img.numel()
                integ.render_backward(scene, params, g, seed=90 + it, spp=spp, shard=shard)
                torch.cuda.synchronize()
            print(f"count {count} block {block:6d} rank {rank}: fwd {scene.ctx.kernel_ms(0):.2f} ms bwd {scene.ctx.kernel_ms(1):.2f} ms", flush=True)

if __name__ == "__main__":
    main(**{k: int(v) for k, v in (a.split("=") for a in sys.argv[1:])})
