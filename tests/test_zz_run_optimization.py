"""GPU: the run around the step -- `run_optimization` (python/optimize.py:275-365) through the drop-in
surface: reference images rendered once and cached as EXR, constant initial grids, per-iteration seeds /
learning rates / upsampling, one random sensor or one ray batch per iteration, loss, backward, optimiser
step + projection, checkpoints (.vol) and previews (.exr)."""
import os

import numpy as np
import pytest

from helpers import hetero_grids

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

KEYS = ["medium1.sigma_t.data", "medium1.albedo.data"]


def _scene_config(uivr, tmp_path, n=16, film=24, **kw):
    sig, alb = hetero_grids(n, seed=5)
    base = dict(volume=uivr.benchmark_scene(n, film, film, scale=6.0, majorant_resolution_factor=8),
                scene_sensors=uivr.circle_sensors(4, film, film), param_keys=list(KEYS), sensors=[0, 1, 3],
                start_from_value={KEYS[0]: 0.3, KEYS[1]: 0.6}, max_depth=8, ref_spp=64,
                references=str(tmp_path / "refs"), ref_params={KEYS[0]: sig, KEYS[1]: alb})
    os.makedirs(base["references"], exist_ok=True)
    base.update(kw)
    return uivr.SceneConfig("run-test", **base)


def test_run_optimization_first_iteration_equals_the_manual_sequence(uivr, tmp_path):
    """Sensor mode, one iteration: the same numbers as spelling optimize.py:325-354 out by hand."""
    dev = torch.device("cuda:0")
    sc = _scene_config(uivr, tmp_path)
    oc = uivr.OptimizationConfig("one", spp=8, n_iter=1, lr=2e-2, primal_spp_factor=2, base_seed=4321,
                                 render_initial=False, render_final=False, checkpoint_initial=False)
    out = str(tmp_path / "out")
    losses = []
    scene, params, opt = uivr.run_optimization(out, oc, sc, "volpathsimple-drt", device=0,
                                               callback=lambda it, l: losses.append(l))
    torch.cuda.synchronize()
    assert len(losses) == 1 and np.isfinite(losses[0]) and set(opt.t.values()) == {1}
    refs = sorted(os.listdir(sc.references))
    assert refs == ["ref_000000.exr", "ref_000001.exr", "ref_000003.exr"]
    assert sorted(os.listdir(os.path.join(out, "params"))) == ["final-medium1_albedo.vol", "final-medium1_sigma_t.vol"]
    assert os.path.isfile(os.path.join(out, "ref_0000.exr"))

    # by hand
    sensor_i = sc.sensors[int(uivr.PCG32(initstate=93483).next_float32() * 3)]
    ref = torch.from_numpy(uivr.read_exr(os.path.join(sc.references, f"ref_{sensor_i:06d}.exr"))).to(dev)
    assert tuple(ref.shape) == (24, 24, 3) and float(ref.mean()) > 0.05
    integ = uivr.get_int_config("volpathsimple-drt").create(max_depth=sc.max_depth)
    scene2, p = uivr.initialize_scene(oc, sc, 0)
    assert tuple(p[KEYS[0]].shape) == (16, 16, 16, 1) and scene2.volume.effective_majorant_factor() == 4
    opt2 = uivr.Adam(lr=oc.lr, params=p)
    opt2.set_learning_rate(oc.learning_rates(sc, 0))
    assert opt2.lr == {KEYS[0]: 2e-2, KEYS[1]: 4e-2}
    q = {k: v.requires_grad_(True) for k, v in p.items()}
    image = uivr.render(scene2, q, integ, sensor=sc.scene_sensors[sensor_i], spp=16, spp_grad=8,
                        seed=uivr.tea32(0, 4321), seed_grad=uivr.tea32(1, 4321))
    loss = uivr.losses.l1(image, ref)
    loss.backward()
    assert abs(float(loss.detach()) - losses[0]) < 1e-6
    grads = {k: v.grad for k, v in q.items()}
    for v in q.values():
        v.requires_grad_(False)
    opt2.step(scene2.ctx, grads, max_density=sc.max_density)
    torch.cuda.synchronize()
    for k in KEYS:
        # the first Adam step moves every touched voxel by ~lr * sign(g): identical except where the
        # order of the atomic gradient sums flips the sign of a gradient that cancels to ~0
        differs = (params[k] - p[k]).abs() > 1e-6
        assert float(differs.float().mean()) < 0.01, k
        written, _, _ = uivr.read_vol(os.path.join(out, "params", "final-" + "_".join(k.split(".")[:-1]) + ".vol"))
        assert np.array_equal(written, params[k].cpu().numpy())
    # the cached reference images are reused, not re-rendered
    stamp = {f: os.path.getmtime(os.path.join(sc.references, f)) for f in refs}
    uivr.get_reference_image_paths(sc, device=0)
    assert stamp == {f: os.path.getmtime(os.path.join(sc.references, f)) for f in refs}


def test_run_optimization_ray_batches_with_upsampling(uivr, tmp_path):
    """Ray-batch mode (the reference's production mode) with one upsampling event, strided checkpoints
    and previews; the loss against the reference views goes down."""
    sc = _scene_config(uivr, tmp_path)
    oc = uivr.OptimizationConfig("batch", spp=4, n_iter=24, lr=2e-2, primal_spp_factor=2, batch_size=1024,
                                 lr_schedule=uivr.Schedule.Last25, upsample=[0.25], checkpoint_stride=8,
                                 preview_stride=10, preview_spp=8)
    out = str(tmp_path / "out")
    losses = []
    scene, params, opt = uivr.run_optimization(out, oc, sc, uivr.get_int_config("volpathsimple-drt"), device=0,
                                               callback=lambda it, l: losses.append(l))
    torch.cuda.synchronize()
    scene.ctx.check_watchdog()
    print("run_optimization ray-batch losses:", [round(l, 5) for l in losses])
    assert len(losses) == 24 and np.all(np.isfinite(losses))
    assert tuple(params[KEYS[0]].shape) == (16, 16, 16, 1) and tuple(params[KEYS[1]].shape) == (16, 16, 16, 3)
    assert scene.volume.res == (16, 16, 16) and scene.volume.effective_majorant_factor() == 4
    s, a = params[KEYS[0]], params[KEYS[1]]
    assert float(s.min()) >= 0.0 and float(s.max()) <= 250.0 and float(a.min()) >= 0.0 and float(a.max()) <= 1.0
    assert set(opt.t.values()) == {24 - 6}                       # the Adam state restarted at the upsampling (iteration 6)
    files = sorted(os.listdir(os.path.join(out, "params")))
    assert files == sorted(f"{p}-medium1_{g}.vol" for p in ("initial", "00000008", "00000016", "final") for g in ("sigma_t", "albedo"))
    first, _, _ = uivr.read_vol(os.path.join(out, "params", "initial-medium1_sigma_t.vol"))
    assert first.shape == (8, 8, 8, 1) and np.all(first == np.float32(0.3))    # coarse start: 16 / 2**1
    previews = sorted(f for f in os.listdir(out) if f.endswith(".exr"))
    assert previews == ["opt_00000010_0000.exr", "opt_00000020_0000.exr", "opt_final_0000.exr", "opt_init_0000.exr", "ref_0000.exr"]
    final = uivr.read_exr(os.path.join(out, "opt_final_0000.exr"))
    assert final.shape == (24, 24, 3) and np.all(np.isfinite(final))
    assert np.mean(losses[-6:]) < np.mean(losses[:6]), losses
