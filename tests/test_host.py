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
                                    "albedo_scatters": 1}, 2, True) == 32 + 96 + 4 + 64 + 192 + 24
    json.dumps(bench.workload_config(2, 128))


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
