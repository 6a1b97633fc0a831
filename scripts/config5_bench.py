"""BASELINE.json config 5: 512^3 sigma_t + albedo grids, 1024 x 1024 x 128 spp, DRT forward + backward,
pixels sharded over 8 ranks, ONE NCCL all-reduce of the 2 GiB [d sigma_t | d albedo] buffer inside the step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/config5_bench.py                  # the real thing (8 GPUs)
    python scripts/config5_bench.py emulate=8     # rank 0's share of the same job on ONE GPU, no collective

Timing: CUDA events around [forward, backward, all-reduce] per step, max over ranks; one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import uivr_b200 as u  # noqa: E402
from importlib import import_module  # noqa: E402

sharding = import_module(u.__name__ + ".sharding")


def main(emulate=0, n=512, w=1024, h=1024, spp=128, steps=3, warmup=2):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    count = emulate if (world == 1 and emulate > 1) else world
    sig, alb = u.synthetic_grids(n)
    params = {"m.sigma_t.data": sig.to(dev), "m.albedo.data": alb.to(dev)}
    del sig, alb
    vol = u.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=8)
    scene = u.Scene(vol, local)
    integ = u.get_int_config("volpathsimple-drt").create(max_depth=64)
    shard = sharding.pixel_shard(rank, count)
    grads = sharding.GradientBuffer(vol.res, dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_step = t_fwd = t_bwd = t_ar = 0.0
    for it in range(warmup + steps):
        seed = 1234 + 2 * it
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        img = integ.render(scene, params, seed=seed, spp=spp, shard=shard)
        ev[1].record()
        g = torch.sign(img - 0.5) / img.numel()           # per-pixel separable loss: local pixels only
        integ.render_backward(scene, params, g, seed=u.tea32(seed, 1), spp=spp, shard=shard, out=grads.views())
        ev[2].record()
        grads.all_reduce()
        ev[3].record()
        torch.cuda.synchronize()
        if it >= warmup:
            t_fwd += ev[0].elapsed_time(ev[1]) / steps
            t_bwd += ev[1].elapsed_time(ev[2]) / steps
            t_ar += ev[2].elapsed_time(ev[3]) / steps
            t_step += ev[0].elapsed_time(ev[3]) / steps
    t = torch.tensor([t_step, t_fwd, t_bwd, t_ar], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0])
        samples = w * h * spp if world > 1 else w * h * spp // max(count, 1)
        print(json.dumps({
            "config": f"5: {n}^3 sigma_t + albedo, {w}x{h}x{spp}spp, DRT fwd+bwd, pixel-sharded x{count}"
                      + ("" if world > 1 else f" -- rank 0's share EMULATED on one GPU, no collective"),
            "n_gpus": world, "ms_per_step": ms, "fwd_ms": float(t[1]), "bwd_ms": float(t[2]), "allreduce_ms": float(t[3]),
            "samples_in_step": samples, "msamples_per_s": samples / ms / 1e3,
            "grad_buffer_bytes": grads.flat.numel() * 4,
            "mem_allocated_gb": torch.cuda.max_memory_allocated() / 1e9}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(**{k: int(v) for k, v in (a.split("=") for a in sys.argv[1:])})
