"""Host-side logic and the C-ABI surface (no GPU): IntegratorConfig / plugin mirror keep the
reference's names, defaults and error behaviour; libuivr.so loads and exports every symbol
declared in include/uivr.h; the product fails loudly without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(uivr):
    hdr = open(os.path.join(ROOT, "include", "uivr.h")).read()
    declared = sorted(set(re.findall(r"\b(uivr_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    assert os.path.exists(uivr._native.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(uivr._native.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"libuivr.so does not export {missing}"
    assert sorted(uivr._native.EXPORTS) == declared, "the ctypes binding must cover the whole header"
    assert lib.uivr_version() >= 100


def test_host_side_integer_helpers_match_the_oracle(uivr, oracle):
    # uivr_tea32 / uivr_alt_seed are plain host functions of the library (no device needed)
    assert uivr.tea32(1234, 1) == 0x38fc4d3a == oracle.tea(1234, 1)[0]
    assert uivr.tea32(0, 0) == 0x5df5f2bf
    rng = np.random.default_rng(0)
    for s in rng.integers(0, 1 << 32, size=50):
        assert uivr._native.lib().uivr_alt_seed(int(s)) == oracle.alt_seed(int(s))
        assert uivr._native.lib().uivr_alt_seed_batch(int(s)) == oracle.alt_seed_batch(int(s))


def test_no_cpu_fallback(uivr):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(uivr.NativeError, match="no CPU fallback"):
        uivr._native.Context(0)
    with pytest.raises(uivr.NativeError, match="no CPU fallback"):
        uivr.Scene(uivr.cube_test_scene(8, 8))


def test_null_and_invalid_arguments_return_status_codes(uivr):
    L = uivr._native.lib()
    assert L.uivr_create(0, None) == -1                      # UIVR_ERR_INVALID
    assert L.uivr_destroy(None) == -1
    assert L.uivr_last_error(None) == b"null context"
    assert L.uivr_set_scene(None, None) == -1
    assert L.uivr_render_forward(None, None, 0, 1, None, None, None, None) == -1


def test_integrator_config_registry_matches_reference(uivr):
    # opt_config.py:123-169
    for name in ["fd-forward", "volpathsimple-drt", "volpathsimple-drt-quadratic", "volpathsimple-basic", "nerf"]:
        assert uivr.get_int_config(name).name == name
    cfg = uivr.get_int_config("volpathsimple-drt")
    assert cfg.pretty_name == "Differential Ratio Tracking"
    assert cfg.params == {"type": "volpathsimple", "use_drt": True, "use_drt_subsampling": True, "use_drt_mis": True}
    fd = uivr.get_int_config("fd-forward")
    assert fd.uses_fd and fd.fd_epsilon == 5e-3 and fd.fd_spp_multiplier == 16
    with pytest.raises(AssertionError):
        uivr.add_int_config("volpathsimple-drt", pretty_name="dup", params={})
    with pytest.raises(AssertionError):
        uivr.IntegratorConfig("x", "x", {}, uses_fd=True)   # fd_epsilon required (opt_config.py:93-95)


def test_integrator_config_create(uivr):
    cfg = uivr.get_int_config("volpathsimple-drt")
    with pytest.raises(AssertionError):
        cfg.create()                                         # max_depth is mandatory (opt_config.py:98)
    with pytest.raises(AssertionError):
        cfg.create(max_depth=8, rr_depth=4)                  # opt_config.py:104
    with pytest.raises(AssertionError):
        cfg.create(max_depth=-1)
    integ = cfg.create(max_depth=8)
    assert isinstance(integ, uivr.VolpathSimpleIntegrator)
    assert (integ.max_depth, integ.rr_depth) == (8, 1008)    # opt_config.py:105-106
    assert (integ.use_nee, integ.use_drt, integ.use_drt_subsampling, integ.use_drt_mis) == (True, True, True, True)
    assert integ.hide_emitters is False and integ.aovs() == []
    basic = uivr.get_int_config("volpathsimple-basic").create(max_depth=3)
    assert basic.use_drt is False
    nerf = uivr.get_int_config("nerf").create(max_depth=3)   # opt_config.py:162-169, nerf.py:27-35
    assert isinstance(nerf, uivr.NeRFIntegrator)
    assert (nerf.queries_per_ray, nerf.jittering_enabled, nerf.activation_type, nerf.hide_emitters) == \
        (128, True, "identity", False)
    assert nerf.rr_depth == 1003 and nerf.aovs() == []
    with pytest.raises(ValueError, match="Unsupported activation"):
        uivr.NeRFIntegrator({"activation": "softplus"}).props()          # nerf.py:44
    with pytest.raises(NotImplementedError):
        uivr.NeRFIntegrator({"density_noise_std": 0.1})                   # nerf.py:157
    with pytest.raises(NotImplementedError):
        uivr.load_dict({"type": "path"})
    # cfg objects handed out are copies (get_int_config deep-copies)
    cfg.params["use_drt"] = False
    assert uivr.get_int_config("volpathsimple-drt").params["use_drt"] is True


def test_plugin_properties_and_errors(uivr):
    d = uivr.VolpathSimpleIntegrator({})                     # defaults of volpathsimple.py:19-34
    assert (d.hide_emitters, d.use_nee, d.use_drt, d.use_drt_subsampling, d.use_drt_mis) == \
        (False, True, True, True, True)
    with pytest.raises(ValueError):
        uivr.VolpathSimpleIntegrator({"no_such_prop": 1})
    with pytest.raises(NotImplementedError):
        uivr.VolpathSimpleIntegrator({"max_depth": 8, "rr_depth": 4})
    assert "volpathsimple" in uivr.INTEGRATORS
    uivr.register_integrator("volpathsimple-alias", lambda props: uivr.VolpathSimpleIntegrator(props))
    assert isinstance(uivr.load_dict({"type": "volpathsimple-alias", "max_depth": 2}), uivr.VolpathSimpleIntegrator)
    with pytest.raises(ValueError):
        uivr.load_dict({"max_depth": 2})


def test_render_seed_rules(uivr):
    import torch
    params = {"m.sigma_t.data": torch.zeros(2, 2, 2, 1), "m.albedo.data": torch.zeros(2, 2, 2, 3)}
    integ = uivr.VolpathSimpleIntegrator({"max_depth": 2})
    # batched.py:122-124: equal primal and differential seeds are an error
    with pytest.raises(Exception, match="seed should be different"):
        uivr.render(None, params, integ, spp=1, seed=5, seed_grad=5)
    # util.get_single_medium (util.py:82-85): exactly one grid of each kind
    with pytest.raises(ValueError):
        uivr.render(None, {"a.sigma_t.data": params["m.sigma_t.data"], "b.sigma_t.data": params["m.sigma_t.data"],
                           "m.albedo.data": params["m.albedo.data"]}, integ, spp=1, seed=5)


def test_scene_description(uivr):
    vol = uivr.cube_test_scene(128, 64, density_scale=2.0)
    d = vol.as_dict()
    # tests/test_integrators.py:40: to_world = translate(-0.5) scale(2) => local = (p + 0.5) / 2
    m = d["to_local"].reshape(3, 4)
    assert np.allclose(m @ np.array([-0.5, -0.5, -0.5, 1.0]), 0) and np.allclose(m @ np.array([1.5, 1.5, 1.5, 1.0]), 1)
    assert d["width"] == 128 and d["height"] == 64 and d["majorant_factor"] == 0
    assert np.isclose(d["tan_x"], np.tan(np.radians(15.0))) and np.isclose(d["tan_y"], d["tan_x"] * 0.5)
    f = np.stack([d["cam_left"], d["cam_up"], d["cam_dir"]])
    assert np.allclose(f @ f.T, np.eye(3), atol=1e-6)
    target = np.array([0.0, -0.15, 0.0]) - np.array([4.0, 4.0, 4.0])
    assert np.allclose(d["cam_dir"], target / np.linalg.norm(target), atol=1e-6)
    # optimize.py:182-199: the supergrid factor is reduced until the supergrid has >= 4 cells per side
    b = uivr.benchmark_scene(256, 8, 8)
    assert b.effective_majorant_factor() == 8
    b16 = uivr.benchmark_scene(16, 8, 8)
    assert b16.effective_majorant_factor() == 4
    b3 = uivr.benchmark_scene(3, 8, 8)
    assert b3.effective_majorant_factor() == 0


def test_synthetic_grids_recipe(uivr):
    sig, alb = uivr.synthetic_grids(32)
    assert tuple(sig.shape) == (32, 32, 32, 1) and tuple(alb.shape) == (32, 32, 32, 3)
    assert float(sig.max()) == 1.0 and float(sig.min()) == 0.0
    assert float((sig == 0).float().mean()) > 0.3            # guaranteed empty space
    assert 0.2 <= float(alb.min()) and float(alb.max()) <= 0.95
    sig2, _ = uivr.synthetic_grids(32)
    assert (sig == sig2).all()


def test_bench_reference_arm_contract():
    import json
    import subprocess
    import sys
    env = dict(os.environ, RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""   # ranks > 0 exit without work
    import bench
    assert bench.algorithmic_bytes({"sigma_taps": 1, "albedo_taps": 1, "majorant_reads": 1, "sigma_scatters": 1,
                                    "albedo_scatters": 1}, 2) == 32 + 96 + 4 + 64 + 192 + 24
    for name in bench.WORKLOADS:
        json.dumps(bench.workload_config(name, 2, 128, "weak"))
    # the parity leg never takes the bench line down: without a device it reports the error instead
    r = bench.parity_check()
    assert set(r) == {"error"}
    # profiler-derived constants are refused when the kernel sources have changed since their capture
    const, why = bench.profiler_constants()
    assert (const is None and why) or const["source_sha"] == bench.kernel_source_sha()


def test_batch_index_sampler_matches_oracle(uivr, oracle):
    """Host restatement of sample_batch_pixels (batched.py:397-423) vs the oracle's per-element
    sampler: same sensors / pixels, bit for bit, and inside the film."""
    sensors = uivr.circle_sensors(5, 40, 24)
    table = uivr.sensor_table(sensors)
    assert table.shape == (5, 16) and table.dtype == np.float32
    for seed in (0, 1234, 0xDEADBEEF):
        si, px = uivr.sample_batch_pixels(300, 5, (40, 24), seed)
        ref = oracle.batch_elements(table, (40, 24), seed, 300)
        assert np.array_equal(si, ref[:, 0]) and np.array_equal(px, ref[:, 1:])
        assert si.max() < 5 and px[:, 0].max() < 40 and px[:, 1].max() < 24
    # every sensor / film region is hit (uniform sampling)
    si, px = uivr.sample_batch_pixels(4000, 5, (40, 24), 7)
    assert set(si.tolist()) == set(range(5)) and px[:, 0].min() == 0 and px[:, 0].max() == 39
    with pytest.raises(ValueError):
        uivr.sensor_table([uivr.Sensor(width=8, height=8), uivr.Sensor(width=9, height=8)])


def test_multires_host_rules(uivr):
    """adjust_majorant_res_factor (optimize.py:182-199) and the upsampling schedule (opt_config.py:40-44)."""
    f = uivr.adjust_majorant_res_factor
    assert f(8, (256, 256, 256, 1)) == 8
    assert f(8, (32, 32, 32, 1)) == 8        # 32 // 8 = 4: still a meaningful supergrid
    assert f(8, (16, 16, 16, 1)) == 4        # largest factor with >= 4 cells
    assert f(8, (6, 6, 6, 1)) == 0           # -> 1 -> supergrid disabled
    assert f(8, (64, 16, 32, 1)) == 4        # shortest axis decides
    assert f(0, (256, 256, 256, 1)) == 0 and f(1, (256,) * 3) == 0
    assert uivr.upsample_iterations([0.04, 0.16, 0.36, 0.64], 6000) == {240, 960, 2160, 3840}  # reproduce.py:58
    assert uivr.upsample_iterations(None, 10) == set()


def test_vol_roundtrip_and_header(uivr, tmp_path):
    """Mitsuba VOL v3 layout (util.save_params -> mi.VolumeGrid.write, util.py:55-71): 48-byte header,
    float32 data with x fastest, then channels."""
    rng = np.random.default_rng(0)
    g = rng.random((5, 4, 3, 3)).astype(np.float32)  # (Z, Y, X, C)
    p = tmp_path / "a.vol"
    uivr.write_vol(str(p), g, (-0.5, -0.5, -0.5), (1.5, 1.5, 1.5))
    raw = p.read_bytes()
    assert raw[:4] == b"VOL\x03" and len(raw) == 48 + g.size * 4
    assert np.frombuffer(raw[4:24], dtype="<i4").tolist() == [1, 3, 4, 5, 3]
    assert np.allclose(np.frombuffer(raw[24:48], dtype="<f4"), [-0.5, -0.5, -0.5, 1.5, 1.5, 1.5])
    # value (z=2, y=1, x=0, c=1) sits at ((z*Y + y)*X + x)*C + c
    off = 48 + 4 * (((2 * 4 + 1) * 3 + 0) * 3 + 1)
    assert np.frombuffer(raw[off:off + 4], dtype="<f4")[0] == g[2, 1, 0, 1]
    back, lo, hi = uivr.read_vol(str(p))
    assert np.array_equal(back, g) and lo == (-0.5, -0.5, -0.5) and hi == (1.5, 1.5, 1.5)
    import torch
    out = uivr.save_params(str(tmp_path / "params"), {"medium1.sigma_t.data": torch.from_numpy(g[..., :1].copy())}, "final")
    assert os.path.basename(out["medium1.sigma_t.data"]) == "final-medium1_sigma_t.vol"
    with pytest.raises(NotImplementedError):
        uivr.save_params(str(tmp_path), {"medium1.scale": torch.zeros(1)}, "x")
    (tmp_path / "bad.vol").write_bytes(b"nope")
    with pytest.raises(ValueError):
        uivr.read_vol(str(tmp_path / "bad.vol"))


# ---- the run around the step: OptimizationConfig / SceneConfig / run_optimization helpers ----------

def _scene_config(uivr, **kw):
    sensors = uivr.circle_sensors(4, 16, 12)
    base = dict(volume=uivr.benchmark_scene(16, 16, 12), scene_sensors=sensors,
                param_keys=["medium1.sigma_t.data", "medium1.albedo.data"], sensors=[0, 2, 3],
                start_from_value={"medium1.sigma_t.data": 0.04, "medium1.albedo.data": 0.6})
    base.update(kw)
    return uivr.SceneConfig("t", **base)


def test_optimization_config_matches_reference(uivr):
    """OptimizationConfig (opt_config.py:11-75): defaults, upsample_at, should_upsample, the learning
    rates against the reference's own OptimizationConfig run by refshim (refshim_host.npz)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "refshim_host.npz"))
    keys = ["m.sigma_t.data", "m.albedo.data", "m.emission.data"]
    import types
    sc = types.SimpleNamespace(param_lr_factors={"m.albedo.data": 2.0}, param_keys=keys)
    for name, sched in (("last25", uivr.Schedule.Last25), ("constant", uivr.Schedule.Constant), ("none", None)):
        cfg = uivr.OptimizationConfig("x", spp=4, n_iter=101, lr=5e-3, lr_schedule=sched)
        got = np.array([[cfg.learning_rates(sc, it)[k] for k in keys] for it in range(101)])
        assert np.array_equal(got, g[f"lr/{name}"])
    with pytest.raises(ValueError, match="Unsupported schedule"):
        uivr.OptimizationConfig("x", spp=4, n_iter=10, lr=1.0, lr_schedule=7).learning_rates(sc, 0)
    for i, (ups, n_iter) in enumerate((([0.25, 0.5, 0.75], 101), ([0.0, 1.0, 0.333], 3000), (None, 10))):
        cfg = uivr.OptimizationConfig("x", spp=4, n_iter=n_iter, lr=1.0, upsample=ups)
        assert sorted(cfg.upsample_at) == list(g[f"upsample_at/{i}"])
        assert np.array_equal(np.array([cfg.should_upsample(it) for it in range(n_iter)]), g[f"should_upsample/{i}"])
    d = uivr.OptimizationConfig("x", spp=4, n_iter=10, lr=1.0)
    # positional construction binds the reference's fields (opt_config.py:14-37, in this order)
    import dataclasses
    assert [f.name for f in dataclasses.fields(uivr.OptimizationConfig)] == [
        "name", "spp", "n_iter", "lr", "primal_spp_factor", "batch_size", "lr_schedule", "upsample", "base_seed",
        "render_initial", "render_final", "preview_stride", "checkpoint_initial", "checkpoint_final", "checkpoint_stride",
        "preview_spp", "opt_type", "opt_args", "loss"]
    pos = uivr.OptimizationConfig("x", 4, 100, 1.0, 64, None, uivr.Schedule.Last25, [0.5])
    assert pos.lr_schedule == uivr.Schedule.Last25 and pos.upsample == [0.5] and pos.base_seed == 988378
    assert (d.primal_spp_factor, d.batch_size, d.base_seed, d.preview_stride, d.checkpoint_stride, d.opt_type) == \
        (64, None, 988378, 100, 1000, "adam")
    assert d.loss is uivr.losses.l1 and d.render_initial and d.render_final and d.checkpoint_initial and d.checkpoint_final
    with pytest.raises(AssertionError):
        uivr.OptimizationConfig("x", spp=4, n_iter=10, lr=1.0, upsample=[1.5])
    with pytest.raises(KeyError):
        uivr.OptimizationConfig("x", spp=4, n_iter=10, lr=1.0, opt_type="lbfgs").optimizer({})


def test_sgd_optimizer_and_projection(uivr):
    import torch
    p = {"m.sigma_t.data": torch.tensor([0.5, 249.0, 0.01]), "m.albedo.data": torch.tensor([0.5, 0.99, 0.0])}
    opt = uivr.OptimizationConfig("x", spp=1, n_iter=2, lr=1.0, opt_type="sgd").optimizer(p)
    assert isinstance(opt, uivr.SGD)
    opt.set_learning_rate({"m.sigma_t.data": 2.0})
    opt.step(None, {"m.sigma_t.data": torch.tensor([0.1, -1.0, 1.0]), "m.albedo.data": torch.tensor([0.25, -0.5, 0.5])})
    assert torch.allclose(p["m.sigma_t.data"], torch.tensor([0.3, 250.0, 0.0]))   # clipped to [0, max_density]
    assert torch.allclose(p["m.albedo.data"], torch.tensor([0.25, 1.0, 0.0]))      # clipped to [0, 1]
    m = uivr.SGD(lr=1.0, params={"m.albedo.data": torch.tensor([0.5])}, momentum=0.5)
    for want in (0.4, 0.25):                                                       # v = 0.1, then 0.5*0.1 + 0.1
        m.step(None, {"m.albedo.data": torch.tensor([0.1])})
        assert abs(float(m.params["m.albedo.data"]) - want) < 1e-6
    m.step(None, {})                                                               # no gradient: parameter untouched
    assert abs(float(m.params["m.albedo.data"]) - 0.25) < 1e-6


def test_scene_config_rules(uivr, tmp_path):
    sc = _scene_config(uivr)
    assert sc.preview_sensors == [0] and sc.param_lr_factors == {"medium1.albedo.data": 2.0}
    assert (sc.max_depth, sc.ref_spp, sc.ref_integrator, sc.max_density, sc.majorant_resolution_factor) == \
        (64, 8192, "volpathsimple", 250, 8)
    assert sc.references.endswith(os.path.join("references", "t")) and sc.ref_volume == sc.volume
    with pytest.raises(ValueError, match="was not given an initial value"):
        _scene_config(uivr, start_from_value={"medium1.sigma_t.data": 0.04})
    with pytest.raises(ValueError, match="not part of the scene"):
        _scene_config(uivr, sensors=[0, 9])
    assert _scene_config(uivr, references=str(tmp_path)).references == str(tmp_path)   # an existing directory is kept
    assert _scene_config(uivr, references="shared").references.endswith(os.path.join("references", "shared"))
    M = __import__("importlib").import_module(uivr.__name__ + ".scene_config")
    name = "cfg-test-base"
    if name not in M._SCENE_CONFIGS:
        kw = {f.name: getattr(sc, f.name) for f in __import__("dataclasses").fields(sc) if f.name in
              ("volume", "scene_sensors", "param_keys", "sensors", "start_from_value")}
        uivr.add_scene_config(name, **kw)
        uivr.add_scene_config_variant(name + "-deep", name, max_depth=128, sensors=[1])
    with pytest.raises(AssertionError, match="Duplicate"):
        uivr.add_scene_config(name, **M._SCENE_CONFIG_KWARGS[name])
    v = uivr.get_scene_config(name + "-deep")
    assert (v.max_depth, v.sensors, v.preview_sensors) == (128, [1], [1]) and uivr.get_scene_config(name).max_depth == 64
    v.sensors.append(2)                                                                 # private copy
    assert uivr.get_scene_config(name + "-deep").sensors == [1]


def test_run_helpers(uivr):
    # PCG32 known answers (pcg32-demo): seed (42, 54)
    r = uivr.PCG32(42, 54)
    assert [r.next_uint32() for _ in range(6)] == [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e]
    r = uivr.PCG32()
    assert [r.next_uint32() for _ in range(3)] == [0x1bbeb4f2, 0xe82e89e9, 0x681cfdeb]
    r, f = uivr.PCG32(initstate=93483), uivr.PCG32(initstate=93483)                    # optimize.py:291
    for _ in range(100):
        u32, x = r.next_uint32(), f.next_float32()
        assert 0.0 <= x < 1.0 and x == float(np.float32((u32 >> 9) * 2.0 ** -23))
    # the scalar generator and the vectorised one of batched.py (seeded through TEA like the `independent`
    # sampler) are the same PCG32
    Bm = __import__("importlib").import_module(uivr.__name__ + ".batched")
    vec = Bm._Pcg32(seed=99, n=5)
    v0, v1 = Bm._tea(np.full(5, 99, dtype=np.uint32), np.arange(5, dtype=np.uint32))
    scal = [uivr.PCG32(initstate=int(a), initseq=int(b)) for a, b in zip(v0, v1)]
    for _ in range(20):
        assert list(vec.next_1d()) == [np.float32(g.next_float32()) for g in scal]
    # optimize.py:36-41
    assert uivr.reference_pass_plan((720, 720), 8192, 720 * 720 * 2048) == (4, 2048)
    assert uivr.reference_pass_plan((720, 620), 8192, 720 * 720 * 2048) == (4, 2048)
    assert uivr.reference_pass_plan((64, 64), 8192, 720 * 720 * 2048) == (1, 8192)
    assert uivr.reference_pass_plan((1000, 1000), 100, 30_000_000) == (4, 25)
    assert uivr.reference_pass_plan((1000, 1000), 100, 33_000_000) == (4, 25)
    assert uivr.reference_pass_plan((1000, 1000), 10, 3_400_000) == (3, 4)              # 3 x 4 >= 10
    # optimize.py:146-156
    assert uivr.initial_resolution((64, 64, 64, 3), None) == (64, 64, 64, 3)
    assert uivr.initial_resolution((64, 48, 32, 1), [0.2, 0.5]) == (16, 12, 8, 1)
    with pytest.raises(ValueError, match="Initial resolution not supported"):
        uivr.initial_resolution((64, 64, 7, 1), [0.1, 0.2])
    # optimize.py:255-268, :110-123
    cfg = uivr.OptimizationConfig("x", spp=1, n_iter=10, lr=1.0, checkpoint_stride=4, checkpoint_final=False, render_initial=False)
    assert [uivr.checkpoint_prefix(cfg, i) for i in (0, 3, 4, 8)] == [None, None, "00000004", "00000008"]
    assert uivr.checkpoint_prefix(cfg, "initial") == "initial" and uivr.checkpoint_prefix(cfg, "final") is None
    cfg.checkpoint_stride = 0
    assert uivr.checkpoint_prefix(cfg, 8) is None
    with pytest.raises(ValueError, match="Unsupported"):
        uivr.checkpoint_prefix(cfg, "halfway")
    assert [uivr.preview_suffix(cfg, i) for i in ("initial", "final", 300, "_x")] == [None, "_final", "_00000300", "_x"]


def test_exr_against_opencv(uivr, tmp_path):
    """write_exr / read_exr against OpenCV's OpenEXR codec (independent implementation)."""
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    f = str(tmp_path / "a.exr")
    try:
        ok = cv2.imwrite(f, np.zeros((2, 2, 3), np.float32)) and cv2.imread(f, cv2.IMREAD_UNCHANGED) is not None
    except cv2.error:
        ok = False
    for shape in ((5, 7, 3), (33, 20, 3), (16, 16, 4), (1, 1, 3), (70, 129, 3)):
        a = (rng.random(shape) * 10 - 2).astype(np.float32)
        bgr = [2, 1, 0] + ([3] if shape[2] == 4 else [])
        for comp in ("NONE", "ZIPS", "ZIP"):
            uivr.write_exr(f, a, comp)
            assert np.array_equal(uivr.read_exr(f), a)
            if ok:
                assert np.array_equal(cv2.imread(f, cv2.IMREAD_UNCHANGED)[..., bgr], a), (shape, comp)
        if ok:
            cv2.imwrite(f, a[..., bgr])
            assert np.array_equal(uivr.read_exr(f), a)
            cv2.imwrite(f, a[..., bgr], [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF])
            assert np.array_equal(uivr.read_exr(f), a.astype(np.float16).astype(np.float32))
            # PIZ (wavelet + Huffman; what mi.Bitmap.write and HDRI libraries produce) is read, not written:
            # noise exercises the 16-bit wavelet path, a smooth image the 14-bit path and the run-length symbol
            yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
            smooth = np.stack([np.sin(xx / 7.0 + k) * np.cos(yy / 5.0) + 1.5 for k in range(shape[2])], -1).astype(np.float32)
            for img in (a, smooth, np.full(shape, 0.25, np.float32)):
                for extra, want in (([], img), ([cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF],
                                                img.astype(np.float16).astype(np.float32))):
                    cv2.imwrite(f, img[..., bgr], [cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_PIZ] + extra)
                    assert np.array_equal(uivr.read_exr(f), want), shape
            cv2.imwrite(f, a[..., bgr], [cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_RLE])
            with pytest.raises(NotImplementedError, match="RLE"):
                uivr.read_exr(f)
    with pytest.raises(NotImplementedError):
        uivr.write_exr(f, a, "PIZ")
    with pytest.raises(ValueError):
        uivr.write_exr(f, np.zeros((4, 4)))
    open(f, "wb").write(b"not an exr file at all")
    with pytest.raises(ValueError, match="not an OpenEXR"):
        uivr.read_exr(f)
    # reference images with an alpha channel load as RGB (optimize.py:75-88 reads whatever the file holds)
    rgba = rng.random((6, 5, 4)).astype(np.float32)
    uivr.write_exr(f, rgba)
    refs = uivr.load_reference_images({3: f, 7: f}, batchify=True)
    assert tuple(refs.shape) == (2, 6, 5, 3) and np.array_equal(refs[1].numpy(), rgba[..., :3])
    assert set(uivr.load_reference_images({3: f})) == {3}


def test_run_optimization_control_flow_with_a_stand_in_renderer(uivr, tmp_path, monkeypatch):
    """The host sequence of run_optimization (optimize.py:275-365) -- which files are written when, seeds,
    sensor choice, upsampling, optimiser calls -- with the device path replaced by a differentiable
    stand-in (no CUDA here; tests/test_zz_run_optimization.py runs the real thing on the GPU)."""
    import importlib
    import torch
    import torch.nn.functional as F
    O = importlib.import_module(uivr.__name__ + ".optimize")
    B = importlib.import_module(uivr.__name__ + ".batched")
    M = importlib.import_module(uivr.__name__ + ".multires")
    I = importlib.import_module(uivr.__name__ + ".integrator")
    calls = []

    class FakeCtx:
        def check_watchdog(self):
            calls.append(("watchdog",))

    class FakeScene:
        def __init__(self, volume, device=None):
            self.volume, self.ctx = volume, FakeCtx()

        def update_medium(self, sigma_t, force=False):
            calls.append(("update_medium", tuple(sigma_t.shape)))

        def update_medium_after_reshape(self, sigma_t):
            calls.append(("reshape", tuple(sigma_t.shape)))

    def shade(params, h, w):
        return (params[KEYS[0]].mean() * params[KEYS[1]].mean(dim=(0, 1, 2))).expand(h, w, 3)

    KEYS = ["medium1.sigma_t.data", "medium1.albedo.data"]

    def fake_integrator_render(self, scene, params, sensor=None, seed=0, spp=0, **kw):
        calls.append(("integrator.render", seed, spp, sensor.width))
        return shade(params, sensor.height, sensor.width).clone()

    def fake_render(scene, params, integrator, sensor=None, spp=0, spp_grad=0, seed=0, seed_grad=0, **kw):
        calls.append(("render", seed, seed_grad, spp, spp_grad, sensor))
        return shade(params, sensor.height, sensor.width)

    def fake_render_batch(batch_size, scene, sensors, params, integrator, seed=0, seed_grad=0, spp=0, spp_grad=0):
        calls.append(("render_batch", batch_size, len(sensors), seed, seed_grad, spp, spp_grad))
        si, px = B.sample_batch_pixels(batch_size, len(sensors), (sensors[0].width, sensors[0].height), seed)
        return shade(params, batch_size, 1).reshape(batch_size, 3), si, px

    monkeypatch.setattr(O, "_cuda_device", lambda device: torch.device("cpu"))
    monkeypatch.setattr(O, "Scene", FakeScene)
    monkeypatch.setattr(O, "render", fake_render)
    monkeypatch.setattr(B, "render_batch", fake_render_batch)
    monkeypatch.setattr(I.VolpathSimpleIntegrator, "render", fake_integrator_render)
    monkeypatch.setattr(M, "upsample_grid", lambda ctx, v, new_res: F.interpolate(
        v.permute(3, 0, 1, 2)[None], size=tuple(new_res[:3]), mode="nearest")[0].permute(1, 2, 3, 0).contiguous())

    sig = np.full((8, 8, 8, 1), 0.7, np.float32)
    alb = np.full((8, 8, 8, 3), 0.9, np.float32)
    sc = _scene_config(uivr, volume=uivr.benchmark_scene(16, 16, 12), references=str(tmp_path / "refs"), ref_spp=100,
                       ref_params={KEYS[0]: sig, KEYS[1]: alb}, max_depth=5, sensors=[0, 2, 3], preview_sensors=[2])
    os.makedirs(sc.references, exist_ok=True)

    # --- sensor mode, SGD, no upsampling
    oc = uivr.OptimizationConfig("s", spp=3, n_iter=5, lr=0.05, primal_spp_factor=4, opt_type="sgd", checkpoint_stride=2,
                                 preview_stride=3, base_seed=77)
    losses = []
    out = str(tmp_path / "sensor")
    scene, params, opt = uivr.run_optimization(out, oc, sc, "volpathsimple-basic", callback=lambda it, l: losses.append(l))
    assert sorted(os.listdir(sc.references)) == ["ref_000000.exr", "ref_000002.exr", "ref_000003.exr"]
    assert np.allclose(uivr.read_exr(os.path.join(sc.references, "ref_000002.exr")), 0.7 * 0.9)
    ref_calls = [c for c in calls if c[0] == "integrator.render" and c[2] == 100]
    assert [c[1] for c in ref_calls] == [1234, 1234, 1234]                       # one pass per sensor at seed 1234
    picks = uivr.PCG32(initstate=93483)
    renders = [c for c in calls if c[0] == "render"]
    assert len(renders) == 5
    for it, c in enumerate(renders):
        want = sc.scene_sensors[sc.sensors[int(picks.next_float32() * 3)]]
        assert c[1:5] == (uivr.tea32(2 * it, 77), uivr.tea32(2 * it + 1, 77), 12, 3) and c[5] is want
    assert losses[-1] < losses[0] and len(losses) == 5                            # sigma_t * albedo moves towards 0.63
    assert sorted(os.listdir(os.path.join(out, "params"))) == sorted(
        f"{p}-medium1_{g}.vol" for p in ("initial", "00000002", "00000004", "final") for g in ("sigma_t", "albedo"))
    assert sorted(f for f in os.listdir(out) if f.endswith(".exr")) == [
        "opt_00000003_0002.exr", "opt_final_0002.exr", "opt_init_0002.exr", "ref_0002.exr"]
    assert [c for c in calls if c[0] == "update_medium"] == [("update_medium", (16, 16, 16, 1))] * 5
    assert calls[-2][0] == "watchdog" and calls[-1][:3] == ("integrator.render", 1234, 3)   # ... then the final preview
    assert tuple(params[KEYS[0]].shape) == (16, 16, 16, 1) and opt.params is params

    # --- ray batches, two upsamplings: 16 -> 4 -> 8 -> 16; the cached references are not rendered again
    calls.clear()
    oc = uivr.OptimizationConfig("b", spp=2, n_iter=8, lr=0.05, batch_size=64, opt_type="sgd", upsample=[0.25, 0.5],
                                 render_initial=False, render_final=False, checkpoint_initial=False, checkpoint_stride=0)
    out = str(tmp_path / "batch")
    scene, params, opt = uivr.run_optimization(out, oc, sc, uivr.get_int_config("volpathsimple-drt"))
    assert not [c for c in calls if c[0] == "integrator.render"]
    rb = [c for c in calls if c[0] == "render_batch"]
    assert [c[1:3] for c in rb] == [(64, 3)] * 8 and all(c[5:] == (128, 2) for c in rb)
    assert [c[3] for c in rb] == [uivr.tea32(2 * it, 988378) for it in range(8)]
    shapes = [c[1][0] for c in calls if c[0] == "update_medium"]
    assert shapes == [4, 4, 8, 8, 16, 16, 16, 16]
    assert [c[1] for c in calls if c[0] == "reshape"] == [(8, 8, 8, 1), (16, 16, 16, 1)]
    # optimize.py:249-250: an upsampling step leaves the medium on the CONFIGURED supergrid factor
    assert scene.volume.res == (16, 16, 16) and scene.volume.majorant_resolution_factor == 8
    assert sorted(os.listdir(os.path.join(out, "params"))) == ["final-medium1_albedo.vol", "final-medium1_sigma_t.vol"]
    assert sorted(f for f in os.listdir(out) if f.endswith(".exr")) == ["ref_0002.exr"]
    with pytest.raises(ValueError, match="Initial resolution not supported"):
        uivr.run_optimization(out, uivr.OptimizationConfig("b", spp=2, n_iter=8, lr=0.05, upsample=[0.1, 0.2, 0.3, 0.4]),
                              sc, "volpathsimple-drt")

    # --- warm start: start_from_value None keeps "what the scene file holds" (optimize.py:141-144), given here as the
    # path of a .vol checkpoint (how the reference chains its nerf results into the next run, scene_config.py)
    warm = np.random.default_rng(2).random((16, 16, 16, 1)).astype(np.float32)
    vol_path = str(tmp_path / "warm-medium1_sigma_t.vol")
    uivr.write_vol(vol_path, warm)
    sc2 = _scene_config(uivr, volume=uivr.benchmark_scene(16, 16, 12), references=sc.references, ref_spp=100,
                        ref_params={KEYS[0]: vol_path, KEYS[1]: alb}, sensors=[0, 2, 3],
                        start_from_value={KEYS[0]: None, KEYS[1]: 0.5}, initial_params={KEYS[0]: vol_path})
    oc = uivr.OptimizationConfig("w", spp=1, n_iter=1, lr=0.0, opt_type="sgd", render_initial=False, render_final=False,
                                 checkpoint_final=False)
    scene, params, opt = uivr.run_optimization(str(tmp_path / "warm"), oc, sc2, "volpathsimple-drt")
    assert np.array_equal(params[KEYS[0]].numpy(), warm) and float(params[KEYS[1]].mean()) == 0.5
    got, _, _ = uivr.read_vol(str(tmp_path / "warm" / "params" / "initial-medium1_sigma_t.vol"))
    assert np.array_equal(got, warm)
    with pytest.raises(AssertionError):                      # a kept grid cannot be combined with upsampling (optimize.py:142)
        uivr.run_optimization(out, uivr.OptimizationConfig("w", spp=1, n_iter=4, lr=0.0, opt_type="sgd", upsample=[0.5]),
                              sc2, "volpathsimple-drt")


def test_exr_roundtrip_properties(uivr, tmp_path):
    """Round trips on arbitrary float bit patterns (NaN payloads, infinities, denormals, -0) and image
    shapes around the 16-scanline ZIP block size; the ZIP byte transform is its own inverse."""
    from hypothesis import given, settings, strategies as st
    E = __import__("importlib").import_module(uivr.__name__ + ".exr")
    f = str(tmp_path / "p.exr")

    @settings(max_examples=40, deadline=None)
    @given(st.integers(1, 40), st.integers(1, 9), st.sampled_from([3, 4]), st.sampled_from(["NONE", "ZIPS", "ZIP"]),
           st.integers(0, 2 ** 32 - 1))
    def images(h, w, c, comp, seed):
        bits = np.random.default_rng(seed).integers(0, 2 ** 32, size=(h, w, c), dtype=np.uint64).astype(np.uint32)
        a = bits.view(np.float32)
        uivr.write_exr(f, a, comp)
        assert np.array_equal(uivr.read_exr(f).view(np.uint32), bits)

    @settings(max_examples=60, deadline=None)
    @given(st.binary(min_size=1, max_size=300))
    def codec(raw):
        assert E._zip_decode(E._zip_encode(raw), len(raw)) == raw

    images()
    codec()

    # damaged files end in ValueError (or decode to something), never in a hang or another exception type
    cv2 = pytest.importorskip("cv2")
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:40, 0:33]
    img = np.stack([np.sin(xx / 5.0 + k) + 1.2 + 0.05 * rng.random((40, 33)) for k in range(3)], -1).astype(np.float32)
    for code in (cv2.IMWRITE_EXR_COMPRESSION_PIZ, cv2.IMWRITE_EXR_COMPRESSION_ZIP):
        try:
            if not cv2.imwrite(f, img, [cv2.IMWRITE_EXR_COMPRESSION, code]):
                continue
        except cv2.error:
            continue
        good = open(f, "rb").read()
        assert uivr.read_exr(f).shape == (40, 33, 3)
        for trial in range(60):
            bad = bytearray(good)
            for _ in range(int(rng.integers(1, 4))):
                bad[int(rng.integers(300, len(bad)))] ^= 1 << int(rng.integers(0, 8))
            open(f, "wb").write(bytes(bad[:len(bad) - int(rng.integers(0, 50)) * (trial % 3 == 0)]))
            try:
                uivr.read_exr(f)
            except (ValueError, NotImplementedError):
                pass


def test_radiance_hdr_against_opencv_and_envmap_from_file(uivr, tmp_path):
    """read_hdr / write_hdr (the reference's envmap format, scene_config.py `envmap_filename: *.hdr`) against
    OpenCV's codec -- run-length encoded and flat scanlines, widths outside the RLE range -- and
    EnvMap.from_file for .hdr / .exr."""
    cv2 = pytest.importorskip("cv2")
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    rng = np.random.default_rng(11)
    f = str(tmp_path / "a.hdr")
    for shape in ((5, 7, 3), (33, 20, 3), (16, 64, 3), (3, 300, 3), (2, 8, 3)):
        a = (rng.random(shape) ** 4 * 50).astype(np.float32)
        a[0, 0] = 0
        a[1, 1] = (1e-6, 2e-6, 0)
        a[:, :shape[1] // 2] = a[:, :1]                       # constant stretches: runs in the RLE stream
        tol = a.max(axis=2, keepdims=True) / 128 + 1e-30      # shared exponent: 8 bits relative to the largest channel
        cv2.imwrite(f, a[..., ::-1])                          # OpenCV writes run-length encoded scanlines
        got = uivr.read_hdr(f)
        assert np.array_equal(got, cv2.imread(f, cv2.IMREAD_UNCHANGED)[..., ::-1]) and np.all(np.abs(got - a) <= tol)
        uivr.write_hdr(f, a)                                  # flat scanlines
        back = uivr.read_hdr(f)
        assert np.array_equal(back, cv2.imread(f, cv2.IMREAD_UNCHANGED)[..., ::-1]) and np.array_equal(back, got)
    env = uivr.EnvMap.from_file(f, scale=1.5)
    assert np.array_equal(env.image, back) and env.scale == 1.5 and env.tables()["env_w"] == 8
    g = str(tmp_path / "a.exr")
    lat = (rng.random((6, 12, 3)) + 0.1).astype(np.float32)
    lat[0, 0] = (-1.0, np.nan, np.inf)                        # bad texels of captured maps are zeroed
    uivr.write_exr(g, lat)
    env = uivr.EnvMap.from_file(g, to_world=((0, 0, 1), (0, 1, 0), (-1, 0, 0)))
    assert np.array_equal(env.image[1:], lat[1:]) and np.all(env.image[0, 0] == 0) and env.tables()["env_h"] == 6
    with pytest.raises(ValueError, match="unsupported environment map format"):
        uivr.EnvMap.from_file(str(tmp_path / "a.png"))
    open(f, "wb").write(b"#?RADIANCE\nFORMAT=32-bit_rle_xyze\n\n-Y 1 +X 1\n\0\0\0\0")
    with pytest.raises(NotImplementedError, match="xyze"):
        uivr.read_hdr(f)
    open(f, "wb").write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n+Y 1 +X 1\n\0\0\0\0")
    with pytest.raises(NotImplementedError, match="orientation"):
        uivr.read_hdr(f)
    open(f, "wb").write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 2 +X 2\n\0\0\0\0")
    with pytest.raises(ValueError, match="truncated"):
        uivr.read_hdr(f)


# ---------------------------------------------------------------------------------------
# bench.py's CPU arm: self-contained workload statement, no product code on that path
# ---------------------------------------------------------------------------------------

def test_oracle_workload_matches_product_workload(uivr):
    """oracle/workload.py restates the bench workloads for the CPU arm (which must not import the product
    package); the two statements must stay equal: grids bit for bit, every field of the scene description."""
    from oracle import workload as W
    for dense in (False, True):
        sig_p, alb_p = uivr.synthetic_grids(24, dense=dense)
        sig_w, alb_w = W.synthetic_grids(24, dense=dense)
        assert np.array_equal(sig_p.numpy(), sig_w) and np.array_equal(alb_p.numpy(), alb_w)
    assert float(uivr.synthetic_grids(24, dense=True)[0].min()) >= 0.5   # no empty space
    for n, w, h in ((256, 512, 512), (512, 1024, 1024), (8, 16, 12)):
        d_p = uivr.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=8).as_dict()
        d_w = W.benchmark_desc(n, w, h)
        assert sorted(d_p) == sorted(d_w)
        for k in d_p:
            assert np.array_equal(np.asarray(d_p[k]), np.asarray(d_w[k])), k
    assert uivr.get_int_config("volpathsimple-drt").create(max_depth=64).props() == W.drt_props()
    from oracle import oracle as O
    assert W.step_seeds(O, 3) == (uivr.tea32(6, 1234), uivr.tea32(7, 1234))
    import bench
    assert {k: v[:4] for k, v in W.WORKLOADS.items()} == {k: v[:4] for k, v in bench.WORKLOADS.items()}


def test_reference_arm_touches_no_product_code():
    """`bench.py --impl reference` and the cpu_baseline leg: no module of the product package imported, libuivr.so
    not mapped (the judge reads the arm's loaded libraries)."""
    import subprocess
    import sys
    code = (
        "import sys, bench\n"
        "O, W, desc, props, sig, alb, shape = bench._cpu_workload('config3')\n"
        "W.oracle_step(O, dict(desc, width=16, height=16), props, sig, alb, 0, 1, 2)\n"
        "mods = [m for m in sys.modules if 'uivr_b200' in m or 'unbiased-inverse' in m]\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert not mods, mods\n"
        "assert 'libuivr.so' not in maps and 'libuivr_oracle' in maps\n"
        "print('clean')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root)
    assert r.returncode == 0 and "clean" in r.stdout, r.stderr[-800:]
