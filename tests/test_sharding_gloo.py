"""N > 1 host logic on CPU: world_size-2 gloo process group, pixel sharding + ONE gradient
all-reduce.  The per-rank renders are done by the CPU oracle here (it implements the same
uivr_shard rule as the kernels); the GPU version of this test lives in test_gpu_parity.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    u = importlib.import_module("uivr_b200")
    sharding = importlib.import_module("uivr_b200.sharding")
    from helpers import FLAG_COMBOS, hetero_grids, loss_grad
    from oracle import oracle as O

    n, w, h, spp = 8, 16, 12, 4
    sig, alb = hetero_grids(n, seed=3)
    vol = u.cube_test_scene(w, h, density_scale=5.0, res=(n, n, n))
    vol.majorant_resolution_factor = 2
    props = dict(max_depth=8, use_nee=True, **FLAG_COMBOS["volpathsimple-drt"])
    shard = sharding.pixel_shard(rank, world, block=5)
    img, _, _ = O.render_forward(vol.as_dict(), props, sig, alb, 3, spp, shard=shard, nthreads=2)
    mask = sharding.owned_pixel_mask(w * h, shard).view(h, w).numpy()
    assert np.all(img[~mask] == 0.0)
    # per-pixel separable loss: the local loss gradient needs local pixels only
    gimg = loss_grad(img) * mask[..., None]
    ds, da, _, _ = O.render_backward(vol.as_dict(), props, sig, alb, gimg, 4, spp, shard=shard, nthreads=2)
    buf = sharding.GradientBuffer(vol.res, "cpu")
    buf.dsigma.copy_(torch.from_numpy(ds).float())
    buf.dalbedo.copy_(torch.from_numpy(da).float())
    buf.all_reduce()
    full = sharding.gather_image(torch.from_numpy(img.copy()))
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), image=full.numpy(), dsigma=buf.dsigma.numpy(),
                 dalbedo=buf.dalbedo.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_pixel_sharded_gradients_allreduce_gloo(tmp_path, uivr, oracle):
    from helpers import FLAG_COMBOS, hetero_grids, loss_grad
    world, port = 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "out.npz")
    n, w, h, spp = 8, 16, 12, 4
    sig, alb = hetero_grids(n, seed=3)
    vol = uivr.cube_test_scene(w, h, density_scale=5.0, res=(n, n, n))
    vol.majorant_resolution_factor = 2
    props = dict(max_depth=8, use_nee=True, **FLAG_COMBOS["volpathsimple-drt"])
    img, _, _ = oracle.render_forward(vol.as_dict(), props, sig, alb, 3, spp)
    ds, da, _, _ = oracle.render_backward(vol.as_dict(), props, sig, alb, loss_grad(img), 4, spp)
    assert np.array_equal(got["image"], img)                  # disjoint pixels: exact
    # fp32 transport of the per-rank partial sums: compare relative to the largest entry
    assert np.max(np.abs(got["dsigma"] - ds)) < 1e-6 * np.max(np.abs(ds))
    assert np.max(np.abs(got["dalbedo"] - da)) < 1e-6 * np.max(np.abs(da))


def test_shard_rule_partitions_pixels(uivr):
    import importlib
    sharding = importlib.import_module("uivr_b200.sharding")
    for world, block, npix in [(2, 64, 1000), (3, 7, 100), (8, 64, 512 * 512), (4, 1, 17)]:
        total = torch.zeros(npix, dtype=torch.int32)
        for r in range(world):
            total += sharding.owned_pixel_mask(npix, sharding.pixel_shard(r, world, block)).int()
        assert torch.all(total == 1)
    assert sharding.pixel_shard(0, 1) is None
    with pytest.raises(ValueError):
        sharding.pixel_shard(2, 2)
    buf = sharding.GradientBuffer((4, 3, 2), "cpu")
    assert tuple(buf.dsigma.shape) == (2, 3, 4, 1) and tuple(buf.dalbedo.shape) == (2, 3, 4, 3)
    buf.dsigma.fill_(1.0)
    buf.dalbedo.fill_(2.0)
    assert float(buf.flat.sum()) == 24 * 1.0 + 72 * 2.0
    assert buf.all_reduce() is None                           # no process group: no-op
