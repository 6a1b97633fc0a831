"""Pinning the CPU oracle (no GPU needed).

The reference holds no golden vectors for this path and cannot be imported here,
so the oracle is pinned by: public integer known answers (KA5), analytic known answers on
homogeneous media (KA1, KA3), finite differences with the reference's own methodology
(KA4, python/fd.py + tests/test_integrators.py:209-218) and the committed golden fixtures.
"""
import math
import os

import numpy as np
import pytest

from helpers import FLAG_COMBOS, hetero_grids, loss_grad, rel_linf

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---------------------------------------------------------------------------------------
# KA5: integer RNG vectors
# ---------------------------------------------------------------------------------------

def test_pcg32_public_known_answers(oracle):
    # pcg32-demo (pcg-c-basic) known answer for seed(42, 54)
    assert [hex(v) for v in oracle.pcg32_stream(42, 54, 6)] == \
        ["0xa15c02b7", "0x7b47f409", "0xba1d3330", "0x83d2f293", "0xbfa4784b", "0xcbed606e"]
    # PCG32 default-constructed stream (pcg32.h PCG32_DEFAULT_STATE / PCG32_DEFAULT_STREAM)
    assert [hex(v) for v in oracle.pcg32_stream(0x853c49e6748fea9b, 0xda3e39cb94b95bdb, 3)] == \
        ["0x1bbeb4f2", "0xe82e89e9", "0x681cfdeb"]


def test_tea_vectors(oracle):
    assert oracle.tea(0, 0) == (0x5df5f2bf, 0x54ce08ba)
    assert oracle.tea(1234, 1) == (0x38fc4d3a, 0x4bd0027c)  # default seed_grad for seed 1234


def test_tea_against_python_restatement(oracle):
    def tea_py(v0, v1, rounds=4):
        s, m = 0, 0xFFFFFFFF
        for _ in range(rounds):
            s = (s + 0x9e3779b9) & m
            v0 = (v0 + ((((v1 << 4) & m) + 0xa341316c) ^ ((v1 + s) & m) ^ ((v1 >> 5) + 0xc8013ea4))) & m
            v1 = (v1 + ((((v0 << 4) & m) + 0xad90777d) ^ ((v0 + s) & m) ^ ((v0 >> 5) + 0x7e95761e))) & m
        return v0, v1
    rng = np.random.default_rng(3)
    for a, b in rng.integers(0, 1 << 32, size=(200, 2)):
        assert oracle.tea(int(a), int(b)) == tea_py(int(a), int(b))


def test_sampler_floats_in_unit_interval_and_independent_streams(oracle):
    a = oracle.sampler_floats(1234, 0, 4096)
    b = oracle.sampler_floats(1234, 1, 4096)
    assert a.min() >= 0.0 and a.max() < 1.0
    assert abs(a.mean() - 0.5) < 0.02 and abs(np.corrcoef(a, b)[0, 1]) < 0.06
    # next_1d = bitcast((u32 >> 9) | 0x3f800000) - 1: multiples of 2^-23
    assert np.all(a * 2 ** 23 == np.round(a * 2 ** 23))


def test_rng_golden(oracle):
    g = np.load(os.path.join(GOLDEN, "rng.npz"))
    assert np.array_equal(oracle.pcg32_stream(42, 54, 6), g["pcg32_42_54"])
    assert np.array_equal(oracle.sampler_floats(1234, 0, 8).view(np.uint32), g["sampler_1234_0"])
    assert oracle.alt_seed(0x38fc4d3a) == int(g["alt_seed_0x38fc4d3a"][0])


# ---------------------------------------------------------------------------------------
# exact-op transcendental replacements
# ---------------------------------------------------------------------------------------

def test_neg_log1m_accuracy(oracle):
    u = np.concatenate([np.linspace(0, 1, 200001, dtype=np.float32)[:-1],
                        np.float32([0.0, 2.0 ** -23, 1 - 2.0 ** -23, 0.5])])
    got = oracle.neg_log1m(u).astype(np.float64)
    # the path computes -log(1 - u) with 1 - u rounded to fp32 first, like the reference's
    # `-dr.log(1 - u)`; the polynomial is judged against the exact log of that fp32 argument
    ref = -np.log((np.float32(1.0) - u).astype(np.float64))
    assert got[-4] == 0.0
    err = np.abs(got - ref) / np.maximum(ref, 1e-30)
    assert np.max(err[ref > 1e-6]) < 4e-7          # ~3 ulp
    assert np.all(np.diff(got[:200000]) >= 0.0)     # monotone => free-flight order preserved


def test_sincos2pi_accuracy(oracle):
    x = np.linspace(0, 1, 100001, dtype=np.float32)[:-1]
    s, c = oracle.sincos2pi(x)
    assert np.max(np.abs(s - np.sin(2 * np.pi * x.astype(np.float64)))) < 5e-7
    assert np.max(np.abs(c - np.cos(2 * np.pi * x.astype(np.float64)))) < 5e-7
    assert np.max(np.abs(s.astype(np.float64) ** 2 + c.astype(np.float64) ** 2 - 1)) < 1e-6


# ---------------------------------------------------------------------------------------
# GridVolume lookup + majorant supergrid (App. B.5, B.8)
# ---------------------------------------------------------------------------------------

def _trilinear_np(grid, p):
    """Independent float64 restatement of the GridVolume lookup (App. B.8)."""
    z, y, x, ch = grid.shape
    res = np.array([x, y, z])
    out = np.zeros((p.shape[0], ch))
    for i, pt in enumerate(p.astype(np.float64)):
        if np.any(pt < 0) or np.any(pt > 1):
            continue
        q = pt * res - 0.5
        i0 = np.floor(q).astype(int)
        w1 = q - i0
        acc = np.zeros(ch)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    xi = min(max(i0[0] + dx, 0), x - 1)
                    yi = min(max(i0[1] + dy, 0), y - 1)
                    zi = min(max(i0[2] + dz, 0), z - 1)
                    w = ((w1[0] if dx else 1 - w1[0]) * (w1[1] if dy else 1 - w1[1]) *
                         (w1[2] if dz else 1 - w1[2]))
                    acc += w * grid[zi, yi, xi]
        out[i] = acc
    return out


@pytest.mark.parametrize("n", [3, 7])
def test_trilinear_against_numpy(oracle, n):
    sig, alb = hetero_grids(n, seed=n)
    rng = np.random.default_rng(0)
    p = rng.random((400, 3)).astype(np.float32)
    p[:8] = [[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    p[8:16] = p[8:16] * 1.3 - 0.15
    assert np.max(np.abs(oracle.trilinear(sig, p) - _trilinear_np(sig, p))) < 2e-6
    assert np.max(np.abs(oracle.trilinear(alb, p) - _trilinear_np(alb, p))) < 2e-6
    # voxel centres reproduce the voxel values exactly
    centres = (np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).reshape(-1, 3)[:, ::-1] + 0.5) / n
    got = oracle.trilinear(sig, centres.astype(np.float32))[:, 0]
    assert np.max(np.abs(got - sig.reshape(-1))) < 1e-6


@pytest.mark.parametrize("n,factor", [(16, 4), (33, 8), (24, 8)])
def test_majorant_bounds_the_density(oracle, n, factor):
    sig, _ = hetero_grids(n, seed=n + 1)
    scale = 3.0
    maj = oracle.build_majorant(sig, scale, factor)
    m = maj.shape[0]
    assert maj.shape == (n // factor,) * 3
    rng = np.random.default_rng(2)
    p = rng.random((20000, 3)).astype(np.float32)
    dens = scale * oracle.trilinear(sig, p)[:, 0]
    cell = np.minimum((p * m).astype(int), m - 1)
    bound = maj[cell[:, 2], cell[:, 1], cell[:, 0]]
    assert np.all(dens <= bound * (1 + 1e-6) + 1e-7)
    assert np.isclose(maj.max(), scale * sig.max())
    # not vacuous: the supergrid is tighter than the global majorant somewhere
    assert maj.min() < maj.max()


def test_exit_mask_against_brute_force(oracle):
    """bit o of a cell <=> every cell of the box between the cell and the grid corner of octant o is empty."""
    rng = np.random.default_rng(5)
    for shape in ((5, 4, 6), (8, 8, 8), (1, 3, 2)):
        maj = (rng.random(shape) * (rng.random(shape) < 0.25)).astype(np.float32)
        mask = oracle.build_exit_mask(maj)
        mz, my, mx = shape
        empty = ~(maj > 0)
        for o in range(8):
            for z in range(mz):
                for y in range(my):
                    for x in range(mx):
                        xs = slice(None, x + 1) if o & 1 else slice(x, None)
                        ys = slice(None, y + 1) if o & 2 else slice(y, None)
                        zs = slice(None, z + 1) if o & 4 else slice(z, None)
                        assert bool((mask[z, y, x] >> o) & 1) == bool(empty[zs, ys, xs].all()), (shape, o, z, y, x)


def test_exit_mask_changes_nothing_but_the_supergrid_reads(oracle, uivr):
    """Stopping a walk once only empty cells are ahead is an optimisation of the supergrid traversal, not
    a change of the estimator: per-sample radiance and gradients are bit-identical with it switched off."""
    n = 16
    sig, alb = hetero_grids(n, seed=3)
    vol = uivr.cube_test_scene(24, 20, density_scale=5.0, res=(n, n, n))
    vol.majorant_resolution_factor = 2
    desc = vol.as_dict()
    props = dict(max_depth=12, use_nee=True, **FLAG_COMBOS["volpathsimple-drt"])
    out = []
    try:
        for on in (True, False):
            oracle.set_exit_mask(on)
            img, smp, cf = oracle.render_forward(desc, props, sig, alb, 11, 4, want_samples=True)
            ds, da, smp_g, cb = oracle.render_backward(desc, props, sig, alb, loss_grad(img), 12, 4, want_samples=True,
                                                       nthreads=1)
            out.append((smp, smp_g, ds, da, cf, cb))
    finally:
        oracle.set_exit_mask(True)
    a, b = out
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])   # one thread: same summation order
    for k in ("sigma_taps", "albedo_taps", "rng_draws", "real_collisions", "sigma_scatters"):
        assert a[4][k] == b[4][k] and a[5][k] == b[5][k]
    # not vacuous: the mask saves supergrid reads on this scene
    assert a[4]["majorant_reads"] < b[4]["majorant_reads"] and a[5]["majorant_reads"] < b[5]["majorant_reads"]


# ---------------------------------------------------------------------------------------
# KA1 / KA3: analytic known answers on homogeneous media
# ---------------------------------------------------------------------------------------

def _chord_through_box(uivr, vol, px, py):
    """World-space chord length of the pixel-centre ray through the medium box (float64)."""
    d = vol.as_dict()
    u_, v_ = (px + 0.5) / d["width"], (py + 0.5) / d["height"]
    cx, cy = float(d["tan_x"]) * (1 - 2 * u_), float(d["tan_y"]) * (1 - 2 * v_)
    dirw = cx * d["cam_left"].astype(np.float64) + cy * d["cam_up"].astype(np.float64) + d["cam_dir"].astype(np.float64)
    dirw /= np.linalg.norm(dirw)
    o = d["cam_origin"].astype(np.float64)
    lo, hi = np.array(vol.bbox_min), np.array(vol.bbox_min) + np.array(vol.bbox_extent)
    t0, t1 = (lo - o) / dirw, (hi - o) / dirw
    tn, tf = np.max(np.minimum(t0, t1)), np.min(np.maximum(t0, t1))
    return max(tf - tn, 0.0) if tf > max(tn, 0) else 0.0


def test_ka1_homogeneous_transmittance_and_its_gradient(uivr, oracle):
    """max_depth = 0: R = Le * 1[no real collision]  =>  E[R] = Le exp(-sigma_t l), and the only
    gradient estimator that fires is the transmittance one on escaped lanes (B5):
    sum_voxels d/d grid = -scale * l * Le exp(-sigma_t l) * dL."""
    n, w, spp = 4, 8, 4096
    grid_val, scale = 0.7, 1.5
    sig = np.full((n, n, n, 1), grid_val, np.float32)
    alb = np.full((n, n, n, 3), 0.6, np.float32)
    vol = uivr.cube_test_scene(w, w, density_scale=scale, res=(n, n, n))
    # (with use_drt the DRT vertex still gathers NEE light regardless of max_depth,
    # volpathsimple.py:621-624, so the closed form below holds for the free-flight estimator only)
    props = dict(max_depth=0, use_nee=True, **FLAG_COMBOS["volpathsimple-basic"])
    img, _, cnt = oracle.render_forward(vol.as_dict(), props, sig, alb, 11, spp)
    le = np.array(vol.radiance)
    sigma_t = grid_val * scale
    worst = 0.0
    for (px, py) in [(4, 4), (3, 5), (2, 2), (0, 0), (6, 3)]:
        ell = _chord_through_box(uivr, vol, px, py)
        # pixel-centre chord vs box-filtered pixel: compare with a loose but meaningful bound
        expect = le * math.exp(-sigma_t * ell)
        if ell == 0.0:
            assert np.allclose(img[py, px], le, atol=1e-6)
        else:
            worst = max(worst, float(np.max(np.abs(img[py, px] - expect) / le)))
    assert worst < 0.05
    assert cnt["albedo_taps"] > 0 and cnt["sigma_scatters"] == 0
    # with DRT enabled the same primal image results (AD flags do not touch the primal)
    img_drt, _, _ = oracle.render_forward(vol.as_dict(), dict(props, **FLAG_COMBOS["volpathsimple-drt"]), sig, alb, 11, spp)
    assert np.array_equal(img, img_drt)
    # gradient: dL = 1 for one pixel only, compare the total against the analytic derivative
    # estimated with the SAME box-filtered jitter by finite differences of the analytic form
    gimg = np.zeros_like(img)
    gimg[4, 4] = 1.0
    ds, da, _, _ = oracle.render_backward(vol.as_dict(), props, sig, alb, gimg, 12, spp)
    assert np.all(da == 0.0)
    # d/d(grid_val) of sum_c pixel = -scale * l * sum_c Le_c exp(-sigma_t l); use the rendered
    # pixel itself for exp(.) and the centre chord for l
    ell = _chord_through_box(uivr, vol, 4, 4)
    analytic = -scale * ell * float(img[4, 4].sum())
    assert abs(ds.sum() - analytic) / abs(analytic) < 0.05


def test_ka3_white_furnace(uivr, oracle):
    """albedo = 1 and unit radiance from everywhere: every path carries radiance 1, so every
    pixel is exactly 1 up to paths cut at max_depth, and all parameter gradients vanish in
    expectation (sharp unbiasedness check of NEE + DRT + MIS together)."""
    n, w, spp = 6, 8, 1024
    sig, _ = hetero_grids(n, seed=5)
    alb = np.ones((n, n, n, 3), np.float32)
    vol = uivr.cube_test_scene(w, w, density_scale=4.0, res=(n, n, n))
    vol.radiance = (1.0, 1.0, 1.0)
    props = dict(max_depth=64, use_nee=True, **FLAG_COMBOS["volpathsimple-drt"])
    img, _, _ = oracle.render_forward(vol.as_dict(), props, sig, alb, 5, spp)
    assert np.max(np.abs(img - 1.0)) < 0.08        # MC noise of the NEE/phase MIS combination
    assert abs(img.mean() - 1.0) < 5e-3
    gimg = np.full_like(img, 1.0 / img.size)
    ds, da, _, _ = oracle.render_backward(vol.as_dict(), props, sig, alb, gimg, 6, spp)
    # reference scale: the same gradient with an absorbing medium (albedo 0.5) is O(1e-1)
    alb2 = np.full_like(alb, 0.5)
    ds2, _, _, _ = oracle.render_backward(vol.as_dict(), props, sig, alb2, gimg, 6, spp)
    assert np.abs(ds.sum()) < 0.03 * np.abs(ds2.sum())


def test_nee_log_capacity_only_moves_event_counts(oracle, uivr):
    """The collision-log model (uivr_oracle_set_nee_log_capacity) is bookkeeping for the parity tests of the CUDA
    adjoint kernel: shadow walks with more tentative collisions than the log holds keep their second walk in the
    main counters; results do not change."""
    rng = np.random.default_rng(5)
    n = 16
    sig = (0.01 + 0.02 * rng.random((n, n, n, 1))).astype(np.float32)
    sig[n // 2, n // 2, n // 2, 0] = 1.0
    alb = (0.3 + 0.6 * rng.random((n, n, n, 3))).astype(np.float32)
    vol = uivr.benchmark_scene(n, 24, 20, scale=60.0, majorant_resolution_factor=16)
    props = dict(max_depth=6)
    img, _, _ = oracle.render_forward(vol.as_dict(), props, sig, alb, 91, 4)
    gimg = loss_grad(img)
    from conftest import NEE_LOG_CAPACITY
    try:
        oracle.set_nee_log_capacity(0)       # unlimited: every second walk is booked as skipped
        ds0, da0, _, cnt0 = oracle.render_backward(vol.as_dict(), props, sig, alb, gimg, 7, 4, want_samples=True)
        rep0 = oracle.last_backward_replay_counters()
        oracle.set_nee_log_capacity(32)
        ds1, da1, _, cnt1 = oracle.render_backward(vol.as_dict(), props, sig, alb, gimg, 7, 4, want_samples=True)
        rep1 = oracle.last_backward_replay_counters()
        over = oracle.nee_log_overflows()
    finally:
        oracle.set_nee_log_capacity(NEE_LOG_CAPACITY)   # the session default (conftest.py)
    assert cnt0 == cnt1 and rel_linf(ds1, ds0) < 1e-5 and rel_linf(da1, da0) < 1e-5
    assert over > 100
    assert 0 < rep1["sigma_taps"] < rep0["sigma_taps"] and rep1["rng_draws"] < rep0["rng_draws"]


def test_radiance_still_to_come_must_be_subtracted_in_path_order(oracle, uivr):
    """Why the CUDA adjoint (which gathers L itself instead of running the reference's primal pass first) keeps the
    reference's ORDER of subtractions (volpathsimple.py:214) when it scatters the vertex gradients afterwards:
    "L - (gathered so far)" is the same quantity with another rounding, harmless on ordinary media (1e-7), but
    the free-flight gradient divides by max(albedo, 1e-8) (:159-161), and where a channel's albedo is exactly 0
    -- the reference's own 3^3 fixture has such a slab -- the radiance still to come is 0 up to rounding and the
    quotient is that rounding noise times 1e8.  Matching the reference there means matching its rounding."""
    out = {}
    for name in ("hetero", "cube3"):
        if name == "cube3":
            sig, alb = uivr.cube_test_grids()
            vol = uivr.cube_test_scene(12, 12, density_scale=2.0)
        else:
            sig, alb = hetero_grids(8, seed=2)
            vol = uivr.cube_test_scene(12, 12, density_scale=4.0, res=(8, 8, 8))
        props = dict(max_depth=8, use_nee=True, **FLAG_COMBOS["volpathsimple-drt"])
        img, _, _ = oracle.render_forward(vol.as_dict(), props, sig, alb, 7, 8)
        res = []
        try:
            for by_difference in (False, True):
                oracle.set_remaining_by_difference(by_difference)
                res.append(oracle.render_backward(vol.as_dict(), props, sig, alb, loss_grad(img), 8, 8, nthreads=1)[:2])
        finally:
            oracle.set_remaining_by_difference(False)
        out[name] = (rel_linf(res[1][0], res[0][0]), rel_linf(res[1][1], res[0][1]))
    assert out["hetero"][0] < 1e-6 and out["hetero"][1] < 1e-6
    assert out["cube3"][0] < 1e-6          # sigma_t: the factor albedo / max(albedo, 1e-8) is 0 there
    assert out["cube3"][1] > 1e-4          # albedo channel with zeros: rounding noise x 1e8
    assert float(uivr.cube_test_grids()[1].min()) == 0.0


# ---------------------------------------------------------------------------------------
# KA2 + an independent primal estimator: checks of the NEE / MIS / ratio-tracking / supergrid arithmetic
# against something that shares no code with the oracle (the role Mitsuba's stock `volpath` plays in the
# reference's test_03, tests/test_integrators.py:222-257)
# ---------------------------------------------------------------------------------------

def _camera_rays(desc, u, v):
    """World-space perspective rays through film positions (u, v) in [0,1]^2 (float64, written from the
    sensor's definition: fov along x, look-at frame), independent of the oracle's camera code."""
    cx = float(desc["tan_x"]) * (1.0 - 2.0 * u)
    cy = float(desc["tan_y"]) * (1.0 - 2.0 * v)
    d = (cx[:, None] * desc["cam_left"].astype(np.float64) + cy[:, None] * desc["cam_up"].astype(np.float64)
         + desc["cam_dir"].astype(np.float64))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.broadcast_to(desc["cam_origin"].astype(np.float64), d.shape)
    return o, d


def _box_span(o, d, lo, hi):
    """[t_near, t_far] of rays against an axis-aligned box; t_far <= max(t_near, 0) means a miss."""
    with np.errstate(divide="ignore", invalid="ignore"):
        t0, t1 = (lo - o) / d, (hi - o) / d
    tn = np.max(np.minimum(t0, t1), axis=1)
    tf = np.min(np.maximum(t0, t1), axis=1)
    return np.maximum(tn, 0.0), tf


def test_ka2_single_scatter_closed_form(uivr, oracle):
    """Homogeneous medium, constant emitter, max_depth = 2: exactly one scattering order carries light.
    E[R] = Le [ exp(-s l) + albedo * s * int_0^l exp(-s t) A(p(t)) dt ],  A(p) = (1/4pi) int exp(-s d(p, w)) dw
    (d = distance from p to the box boundary along w): the NEE half (ratio-tracked transmittance, MIS 1/2) and
    the phase-sampled half (escape after the vertex, MIS 1/2) must add up to the single-scatter integral, which
    is evaluated here by plain quadrature in float64."""
    n, w, spp = 4, 6, 8192
    grid_val, scale, albedo = 0.8, 1.25, 0.7
    s_t = grid_val * scale
    sig = np.full((n, n, n, 1), grid_val, np.float32)
    alb = np.full((n, n, n, 3), albedo, np.float32)
    vol = uivr.cube_test_scene(w, w, density_scale=scale, res=(n, n, n))
    desc = vol.as_dict()
    props = dict(max_depth=2, use_nee=True, **FLAG_COMBOS["volpathsimple-basic"])
    img, _, _ = oracle.render_forward(desc, props, sig, alb, 21, spp)
    lo, hi = np.array(vol.bbox_min, dtype=np.float64), np.array(vol.bbox_min) + np.array(vol.bbox_extent)
    # sphere quadrature: Gauss-Legendre in cos(theta) x uniform in phi
    mu, wmu = np.polynomial.legendre.leggauss(24)
    phi = (np.arange(48) + 0.5) * (2 * np.pi / 48)
    sin_t = np.sqrt(1 - mu ** 2)
    dirs = np.stack([np.outer(sin_t, np.cos(phi)).ravel(), np.outer(sin_t, np.sin(phi)).ravel(),
                     np.repeat(mu, 48)], axis=1)
    wdir = np.repeat(wmu, 48) * (2 * np.pi / 48) / (4 * np.pi)          # sums to 1

    def sphere_mean_transmittance(p):
        _, tf = _box_span(np.broadcast_to(p, dirs.shape), dirs, lo, hi)
        return float(np.sum(wdir * np.exp(-s_t * tf)))

    le = np.array(vol.radiance, dtype=np.float64)
    tq, wq = np.polynomial.legendre.leggauss(32)
    sub = (np.arange(3) + 0.5) / 3                                       # 3 x 3 positions per pixel (box filter)
    worst = 0.0
    for (px, py) in [(2, 3), (3, 2), (1, 1), (4, 4)]:
        acc = 0.0
        for jy in sub:
            for jx in sub:
                o, d = _camera_rays(desc, np.array([(px + jx) / w]), np.array([(py + jy) / w]))
                tn, tf = _box_span(o, d, lo, hi)
                ell = max(float(tf[0] - tn[0]), 0.0) if tf[0] > tn[0] else 0.0
                val = math.exp(-s_t * ell)
                if ell > 0:
                    ts = 0.5 * ell * (tq + 1.0)
                    inner = sum(wk * math.exp(-s_t * tk) * sphere_mean_transmittance(o[0] + (tn[0] + tk) * d[0])
                                for tk, wk in zip(ts, wq)) * 0.5 * ell
                    val += albedo * s_t * inner
                acc += val / 9.0
        expect = le * acc
        worst = max(worst, float(np.max(np.abs(img[py, px] - expect) / le)))
    # per-pixel Monte-Carlo error at 8192 spp: ~0.5 / sqrt(8192) = 0.006 (+ the 3x3 pixel quadrature)
    assert worst < 0.025, worst
    # not vacuous: the single-scatter term is a large part of these pixels
    assert albedo * s_t > 0.5 and float(img.mean()) > 0.3


def _naive_trilinear(grid, p):
    """Trilinear lookup of a (Z,Y,X) grid at local points p in [0,1]^3 (voxel centres at (i + 0.5) / res, border
    clamp) written directly from GridVolume's definition, float64."""
    z, y, x = grid.shape
    q = p * np.array([x, y, z]) - 0.5
    i0 = np.floor(q).astype(int)
    f = q - i0
    out = np.zeros(p.shape[0])
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                ix = np.clip(i0[:, 0] + dx, 0, x - 1)
                iy = np.clip(i0[:, 1] + dy, 0, y - 1)
                iz = np.clip(i0[:, 2] + dz, 0, z - 1)
                wgt = (f[:, 0] if dx else 1 - f[:, 0]) * (f[:, 1] if dy else 1 - f[:, 1]) * (f[:, 2] if dz else 1 - f[:, 2])
                out += wgt * grid[iz, iy, ix]
    return out


def test_primal_against_an_independent_naive_estimator(uivr, oracle):
    """The oracle's primal image (supergrid DDA + exit mask, NEE with ratio tracking, MIS) against a naive
    estimator that shares none of that: numpy, float64, pure delta tracking against ONE global majorant, no
    next-event estimation (light is only picked up when a path escapes), its own trilinear lookup, camera and
    box intersection, numpy's random generator.  Same heterogeneous medium with empty space; both are unbiased
    for the same max_depth-truncated transport, so the images agree to Monte-Carlo error."""
    n, w, h, max_depth = 12, 10, 8, 12
    scale = 6.0
    sig, alb = hetero_grids(n, seed=4)
    vol = uivr.cube_test_scene(w, h, density_scale=scale, res=(n, n, n))
    vol.majorant_resolution_factor = 3
    desc = vol.as_dict()
    assert desc["majorant_factor"] == 3
    props = dict(max_depth=max_depth, use_nee=True, **FLAG_COMBOS["volpathsimple-drt"])
    img_o, _, _ = oracle.render_forward(desc, props, sig, alb, 77, 2048)

    spp = 3000
    rng = np.random.default_rng(123)
    lo = np.array(vol.bbox_min, dtype=np.float64)
    ext = np.array(vol.bbox_extent, dtype=np.float64)
    grid = sig[..., 0].astype(np.float64) * scale
    sig_bar = grid.max()
    le = np.array(vol.radiance, dtype=np.float64)
    pix = np.repeat(np.arange(w * h), spp)
    N = pix.size
    o, d = _camera_rays(desc, ((pix % w) + rng.random(N)) / w, ((pix // w) + rng.random(N)) / h)
    tn, tf = _box_span(o, d, lo, lo + ext)
    hit = tf > tn
    beta = np.ones((N, 3))
    R = np.zeros((N, 3))
    R[~hit] = le
    pos = o + (tn + 1e-9)[:, None] * d
    alive = hit.copy()
    depth = np.zeros(N, dtype=int)
    for _ in range(100000):
        idx = np.nonzero(alive)[0]
        if idx.size == 0:
            break
        t = -np.log1p(-rng.random(idx.size)) / sig_bar
        p_new = pos[idx] + t[:, None] * d[idx]
        local = (p_new - lo) / ext
        inside = np.all((local >= 0) & (local <= 1), axis=1)
        esc = idx[~inside]
        R[esc] += beta[esc] * le                       # the emitter is seen through the boundary
        alive[esc] = False
        idx, p_new, local = idx[inside], p_new[inside], local[inside]
        pos[idx] = p_new
        real = rng.random(idx.size) < _naive_trilinear(grid, local) / sig_bar
        ridx, rloc = idx[real], local[real]
        a = np.stack([_naive_trilinear(alb[..., c].astype(np.float64), rloc) for c in range(3)], axis=1)
        beta[ridx] *= a
        depth[ridx] += 1
        dead = depth[ridx] >= max_depth                # volpathsimple.py:199-200: the path ends, nothing is added
        alive[ridx[dead]] = False
        zc = 1 - 2 * rng.random(ridx.size)
        ph = 2 * np.pi * rng.random(ridx.size)
        rr = np.sqrt(np.maximum(0, 1 - zc * zc))
        d[ridx] = np.stack([rr * np.cos(ph), rr * np.sin(ph), zc], axis=1)
    img_n = R.reshape(h, w, spp, 3).mean(axis=2)
    var_n = R.reshape(h, w, spp, 3).var(axis=2) / spp
    # per pixel: within 5 sigma of the combined Monte-Carlo error (the oracle's NEE estimator has the smaller one)
    err = np.abs(img_n - img_o)
    sigma = np.sqrt(var_n + var_n * spp / 2048.0) + 1e-3
    assert np.all(err < 5 * sigma), float(np.max(err / sigma))
    # whole image: relative difference of the means well below a percent
    assert abs(img_n.mean() - img_o.mean()) / img_o.mean() < 6e-3
    # not vacuous: the medium matters (image far from the bare emitter) and is heterogeneous
    assert float(np.abs(img_o - le).max()) > 0.2 and float((sig == 0).mean()) > 0.2


# ---------------------------------------------------------------------------------------
# KA4: finite differences vs adjoint on the reference's 3^3 fixture
# ---------------------------------------------------------------------------------------

SIG_IDX = [(0, 0, 0), (0, 2, 0), (1, 1, 1), (2, 2, 2)]
ALB_IDX = [(0, 2, 0), (1, 1, 1)]


def _fd_gradients(oracle, desc, props, sig, alb, seed, spp, eps):
    """python/fd.py:9-69: forward differences of the scalar loss, one re-render per entry, same seed."""
    def loss(a_sig, a_alb):
        img, _, _ = oracle.render_forward(desc, props, a_sig, a_alb, seed, spp)
        return float(np.mean((img.astype(np.float64) - 0.5) ** 2))  # tests/test_integrators.py:119-120
    l0 = loss(sig, alb)
    out = {}
    for idx in SIG_IDX:
        s2 = sig.copy()
        s2[idx + (0,)] += eps
        out[("sigma",) + idx] = (loss(s2, alb) - l0) / eps
    for idx in ALB_IDX:
        for c in range(3):
            a2 = alb.copy()
            a2[idx + (c,)] += eps
            out[("albedo",) + idx + (c,)] = (loss(sig, a2) - l0) / eps
    return out


@pytest.fixture(scope="module")
def fd_reference(uivr, oracle):
    sig, alb = uivr.cube_test_grids()
    grey = np.repeat(alb[..., 1:2], 3, axis=-1).copy()
    vol = uivr.cube_test_scene(8, 8, density_scale=2.0)  # density_scale=2 as tests:228, :265
    desc = vol.as_dict()
    props = dict(max_depth=8, use_nee=True, use_drt=False)   # the primal does not depend on the AD flags
    spp_fd, eps = 65536, 2e-2
    return {"desc": desc, "sig": sig,
            "colour": (alb, _fd_gradients(oracle, desc, props, sig, alb, 99, spp_fd, eps)),
            "grey": (grey, _fd_gradients(oracle, desc, props, sig, grey, 99, spp_fd, eps))}


# The reservoir of DRT depth subsampling picks a vertex with probability mean_c(w_c / wsum_c) but
# weights it with mean_c(wsum) w / mean_c(w) (volpathsimple.py:751, :757-759): exact only when
# the throughput is grey.  The oracle reproduces that behaviour verbatim, so the subsampled
# combos are checked for unbiasedness with a grey albedo and the others with the RGB fixture.
@pytest.mark.parametrize("combo,albedo_kind", [
    ("volpathsimple-basic", "colour"), ("volpathsimple-drt-quadratic", "colour"),
    ("volpathsimple-drt", "grey"), ("test04-nomis", "grey")])
def test_ka4_fd_vs_adjoint(oracle, fd_reference, combo, albedo_kind):
    desc, sig = fd_reference["desc"], fd_reference["sig"]
    alb, fd = fd_reference[albedo_kind]
    props = dict(max_depth=8, use_nee=True, **FLAG_COMBOS[combo])
    spp = 16384
    img, _, _ = oracle.render_forward(desc, props, sig, alb, 1234, spp)
    ds, da, _, _ = oracle.render_backward(desc, props, sig, alb, loss_grad(img), 777, spp)
    ratios = []
    for key, ref in fd.items():
        got = ds[key[1:] + (0,)] if key[0] == "sigma" else da[key[1:]]
        ratios.append(got / ref)
        assert np.sign(got) == np.sign(ref), (key, got, ref)
    ratios = np.array(ratios)
    # thresholds in the spirit of tests/test_integrators.py:209-218 (rtol 3e-2 for most entries,
    # a hard cap for all); both sides are Monte-Carlo estimates
    assert np.all(np.abs(ratios - 1.0) < 0.25), ratios
    assert np.sum(np.abs(ratios - 1.0) > 0.06) <= 3, ratios
    assert abs(np.median(ratios) - 1.0) < 0.03, ratios


def test_reservoir_channel_bias_is_the_references(oracle, fd_reference):
    """Documented quirk: with RGB-varying throughput the subsampled DRT estimator deviates
    from FD in a channel-dependent way (red/green high, blue low on this fixture) -- a property of
    volpathsimple.py:751-759, reproduced verbatim."""
    desc, sig = fd_reference["desc"], fd_reference["sig"]
    alb, fd = fd_reference["colour"]
    props = dict(max_depth=8, use_nee=True, **FLAG_COMBOS["test04-nomis"])
    img, _, _ = oracle.render_forward(desc, props, sig, alb, 1234, 16384)
    _, da, _, _ = oracle.render_backward(desc, props, sig, alb, loss_grad(img), 777, 16384)
    r = np.array([[da[idx + (c,)] / fd[("albedo",) + idx + (c,)] for c in range(3)] for idx in ALB_IDX])
    assert np.all(r[:, 0] > 1.04) and np.all(r[:, 2] < 0.98), r


# ---------------------------------------------------------------------------------------
# golden fixtures + structural properties
# ---------------------------------------------------------------------------------------

@pytest.mark.parametrize("case", ["hetero12", "cube3"])
def test_oracle_reproduces_golden(oracle, case):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = np.load(os.path.join(GOLDEN, f"{case}.npz"))
    sig, alb, vol, spp, max_depth, seed, seed_grad = mg.case_inputs(case)
    assert np.array_equal(sig, g["sigma_t"]) and np.array_equal(alb, g["albedo"])
    for combo, flags in FLAG_COMBOS.items():
        props = dict(max_depth=max_depth, use_nee=True, **flags)
        img, samples, cf = oracle.render_forward(vol.as_dict(), props, sig, alb, seed, spp, want_samples=True)
        assert np.array_equal(samples.view(np.uint32), g[f"{combo}/samples"])
        assert np.array_equal(img, g[f"{combo}/image"])
        assert [cf[k] for k in oracle.COUNTER_NAMES] == list(g[f"{combo}/counters_fwd"])
        ds, da, sg, cb = oracle.render_backward(vol.as_dict(), props, sig, alb, loss_grad(img), seed_grad, spp,
                                                want_samples=True)
        assert np.array_equal(sg.view(np.uint32), g[f"{combo}/samples_grad_pass"])
        assert [cb[k] for k in oracle.COUNTER_NAMES] == list(g[f"{combo}/counters_bwd"])
        # multi-threaded double accumulation: summation order differs at the 1e-16 level only
        assert np.allclose(ds, g[f"{combo}/dsigma"], rtol=1e-12, atol=1e-18)
        assert np.allclose(da, g[f"{combo}/dalbedo"], rtol=1e-12, atol=1e-18)


def test_sharded_oracle_equals_unsharded(uivr, oracle):
    sig, alb = hetero_grids(8, seed=3)
    vol = uivr.cube_test_scene(12, 10, density_scale=5.0, res=(8, 8, 8))
    vol.majorant_resolution_factor = 2
    props = dict(max_depth=8, use_nee=True, **FLAG_COMBOS["volpathsimple-drt"])
    img, samples, _ = oracle.render_forward(vol.as_dict(), props, sig, alb, 3, 4, want_samples=True)
    gimg = loss_grad(img)
    ds, da, _, _ = oracle.render_backward(vol.as_dict(), props, sig, alb, gimg, 4, 4)
    acc_img, acc_ds, acc_da = np.zeros_like(img), np.zeros_like(ds), np.zeros_like(da)
    for r in range(3):
        i, _, _ = oracle.render_forward(vol.as_dict(), props, sig, alb, 3, 4, shard=(r, 3, 7))
        assert np.all((i == 0) | (i == img))
        acc_img += i
        d1, d2, _, _ = oracle.render_backward(vol.as_dict(), props, sig, alb, gimg, 4, 4, shard=(r, 3, 7))
        acc_ds += d1
        acc_da += d2
    assert np.array_equal(acc_img, img)
    assert np.allclose(acc_ds, ds, rtol=1e-12, atol=1e-18) and np.allclose(acc_da, da, rtol=1e-12, atol=1e-18)


def test_forward_ignores_ad_flags_and_seeds_matter(uivr, oracle):
    sig, alb = uivr.cube_test_grids()
    vol = uivr.cube_test_scene(8, 8, density_scale=2.0)
    base = dict(max_depth=8, use_nee=True)
    a, _, _ = oracle.render_forward(vol.as_dict(), dict(base, **FLAG_COMBOS["volpathsimple-drt"]), sig, alb, 1, 4)
    b, _, _ = oracle.render_forward(vol.as_dict(), dict(base, **FLAG_COMBOS["volpathsimple-basic"]), sig, alb, 1, 4)
    c, _, _ = oracle.render_forward(vol.as_dict(), dict(base, **FLAG_COMBOS["volpathsimple-basic"]), sig, alb, 2, 4)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    # hide_emitters removes directly visible radiance only (volpathsimple.py:268-269)
    h, _, _ = oracle.render_forward(vol.as_dict(), dict(base, use_drt=False, hide_emitters=True), sig, alb, 1, 4)
    assert np.all(h <= a + 1e-7) and h[0, 0].sum() == 0.0 and a[0, 0].sum() > 0.0


# ---------------------------------------------------------------------------------------
# optimiser step: mi.ad.Adam + enforce_valid_params
# ---------------------------------------------------------------------------------------

def test_adam_step_against_float64_restatement(oracle):
    """SURVEY App. B.10 in float64 numpy vs the oracle's float32 restatement, three steps, with
    the projection of optimize.py:169-179 active on part of the data."""
    rng = np.random.default_rng(5)
    n = 1003
    p = rng.uniform(-0.2, 1.2, n).astype(np.float32)
    m = np.zeros(n, np.float32)
    v = np.zeros(n, np.float32)
    p64, m64, v64 = p.astype(np.float64), m.astype(np.float64), v.astype(np.float64)
    lr, b1, b2, eps = 5e-3, 0.9, 0.999, 1e-8
    for t in (1, 2, 3):
        g = rng.normal(0, 1e-3, n).astype(np.float32)
        oracle.adam_step(p, g, m, v, lr, b1, b2, eps, t, 0.0, 1.0)
        m64 = b1 * m64 + (1 - b1) * g
        v64 = b2 * v64 + (1 - b2) * g.astype(np.float64) ** 2
        p64 = np.clip(p64 - lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t) * m64 / (np.sqrt(v64) + eps), 0.0, 1.0)
        assert np.max(np.abs(p - p64)) < 2e-6
        assert p.min() >= 0.0 and p.max() <= 1.0
    assert np.allclose(m, m64, rtol=1e-5, atol=1e-9) and np.allclose(v, v64, rtol=1e-5, atol=1e-12)


def test_learning_rate_schedule_and_bounds(uivr):
    """opt_config.py:50-69 (Last25 halves the rate at 75 %, 85 %, 95 % of the run; per-key factors)
    and optimize.py:169-179."""
    keys = ["m.sigma_t.data", "m.albedo.data"]
    f = {"m.albedo.data": 2.0}
    assert uivr.learning_rates(5e-3, keys, 0, 101, "last25", f) == {keys[0]: 5e-3, keys[1]: 1e-2}
    assert uivr.learning_rates(5e-3, keys, 75, 101, "last25", f)[keys[0]] == 2.5e-3
    assert uivr.learning_rates(5e-3, keys, 85, 101, "last25", f)[keys[0]] == 1.25e-3
    assert uivr.learning_rates(5e-3, keys, 100, 101, "last25", f)[keys[1]] == 1.25e-3
    assert uivr.learning_rates(5e-3, keys, 100, 101, None)[keys[0]] == 5e-3
    with pytest.raises(ValueError):
        uivr.learning_rates(5e-3, keys, 1, 10, "cosine")
    assert uivr.param_bounds("x.sigma_t.data", 250) == (0.0, 250.0)
    assert uivr.param_bounds("x.albedo.data") == (0.0, 1.0)
    with pytest.raises(ValueError):
        uivr.param_bounds("x.bogus")


# ---------------------------------------------------------------------------------------
# nerf integrator (python/integrators/nerf.py)
# ---------------------------------------------------------------------------------------

def test_exp_exact_accuracy(oracle):
    x = np.concatenate([np.linspace(-90.0, 89.0, 100001), -np.random.default_rng(0).random(50000) * 8.0]).astype(np.float32)
    ref = np.exp(np.clip(x.astype(np.float64), -87.0, 88.0))
    assert np.max(np.abs(oracle.exp_exact(x) - ref) / ref) < 2e-7
    assert oracle.exp_exact(np.zeros(1, np.float32))[0] == 1.0


def test_nerf_emission_equal_to_background_is_invariant(uivr, oracle):
    """Known answer: with emission == emitter radiance everywhere, E w + (1 - w) Le == Le for every
    ray whatever sigma_t is (nerf.py:109-143), so the image is flat and d loss / d sigma_t == 0."""
    n = 10
    sig, _ = hetero_grids(n, seed=3)
    vol = uivr.cube_test_scene(20, 16, density_scale=7.0, res=(n, n, n))
    le = np.asarray(vol.radiance, dtype=np.float32)
    em = np.broadcast_to(le, (n, n, n, 3)).copy()
    props = dict(queries_per_ray=48)
    img, smp, cnt = oracle.nerf_forward(vol.as_dict(), props, sig, em, 11, 4, want_samples=True)
    assert np.max(np.abs(smp - le)) < 2e-6
    assert cnt["sigma_taps"] == cnt["albedo_taps"] == 48 * cnt["camera_hits"] > 0
    ds, de, _, _ = oracle.nerf_backward(vol.as_dict(), props, sig, em, np.ones_like(img), 12, 4)
    assert np.abs(ds).max() < 1e-6 * np.abs(de).max()
    # d image / d emission sums to the mean opacity: sum_k w_k over all rays / (spp)
    assert de.sum() > 0


@pytest.mark.parametrize("props,offset", [(dict(queries_per_ray=32), 0.0),
                                          (dict(queries_per_ray=9, jittering_enabled=False), 0.0),
                                          (dict(queries_per_ray=32, activation="relu"), -0.15)])
def test_nerf_adjoint_matches_finite_differences(uivr, oracle, props, offset):
    """The ray marcher is deterministic given the seed, so central differences of the same-seed loss
    check the restated adjoint (path replay + per-step backward_from, nerf.py:109-124) directly."""
    n = 8
    sig, em = hetero_grids(n, seed=n)
    sig = (sig + np.float32(offset)).astype(np.float32)
    vol = uivr.cube_test_scene(12, 10, density_scale=5.0, res=(n, n, n))
    desc = vol.as_dict()
    img, _, _ = oracle.nerf_forward(desc, props, sig, em, 21, 4)
    ds, de, _, _ = oracle.nerf_backward(desc, props, sig, em, loss_grad(img), 21, 4)

    def loss(s, e):
        im, _, _ = oracle.nerf_forward(desc, props, s, e, 21, 4)
        return float(np.mean((im.astype(np.float64) - 0.5) ** 2))

    rng = np.random.default_rng(1)
    for _ in range(3):
        d1 = rng.standard_normal(sig.shape).astype(np.float32)
        d2 = rng.standard_normal(em.shape).astype(np.float32)
        if offset:  # keep away from the relu kink
            d1[np.abs(sig) < 0.02] = 0.0
        eps = 2e-3
        fd = (loss(sig + eps * d1, em + eps * d2) - loss(sig - eps * d1, em - eps * d2)) / (2 * eps)
        an = float((ds * d1).sum() + (de * d2).sum())
        assert abs(fd - an) < 2e-2 * max(abs(fd), abs(an)) + 1e-7


# ---------------------------------------------------------------------------------------
# envmap emitter
# ---------------------------------------------------------------------------------------

def test_atan2_turns_accuracy(oracle):
    import ctypes as C
    rng = np.random.default_rng(0)
    y = rng.standard_normal(200000).astype(np.float32)
    x = rng.standard_normal(200000).astype(np.float32)
    y[:8] = [0, 0, 1, -1, 1, -1, 0.0, 1e-30]
    x[:8] = [1, -1, 0, 0, 1, -1, 0.0, 1.0]
    out = np.zeros_like(y)
    fp = C.POINTER(C.c_float)
    oracle.lib().uivr_oracle_atan2_turns(y.ctypes.data_as(fp), x.ctypes.data_as(fp), y.size, out.ctypes.data_as(fp))
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64)) / (2 * np.pi)
    assert np.max(np.abs(out - ref)) < 1.5e-7
    assert out[0] == 0.0 and out[1] == 0.5 and out[2] == 0.25 and out[3] == -0.25 and out[6] == 0.0


def _envmap_scene(uivr, n=10, w=24, h=20):
    import importlib
    S = importlib.import_module(uivr.__name__ + ".scene")
    rng = np.random.default_rng(5)
    img = (rng.random((12, 20, 3)) ** 2).astype(np.float32)
    img[3, 4] = (40.0, 30.0, 20.0)
    sig, alb = hetero_grids(n, seed=n)
    vol = uivr.cube_test_scene(w, h, density_scale=5.0, res=(n, n, n))
    vol.envmap = S.EnvMap(img, scale=0.7, to_world=((0, 0, 1), (0, 1, 0), (-1, 0, 0)))
    return sig, alb, vol, img


def test_envmap_tables_and_sampling_are_consistent(uivr, oracle):
    """The host-built tables: density integrates to 1; a sampled direction evaluates back (through
    the atan2/acos path) to the radiance and the solid-angle pdf it was sampled with."""
    from oracle import refshim as R
    import ctypes as C
    sig, alb, vol, img = _envmap_scene(uivr)
    desc = vol.as_dict()
    pdf_uv = desc["env_data"][:-1, :-1, 3].astype(np.float64)
    assert abs(pdf_uv.mean() - 1.0) < 1e-6
    assert np.all(np.diff(desc["env_marg"]) >= 0) and desc["env_marg"][-1] == 1.0
    assert np.all(np.diff(desc["env_cond"], axis=1) >= 0) and np.all(desc["env_cond"][:, -1] == 1.0)
    rng = np.random.default_rng(1)
    n = 20000
    xi1, xi2 = rng.random(n).astype(np.float32), rng.random(n).astype(np.float32)
    with R._Session(desc, sig, alb) as S:
        fp = C.POINTER(C.c_float)
        d, pdf, le = np.empty((n, 3), np.float32), np.empty(n, np.float32), np.empty((n, 3), np.float32)
        R._lib().uivr_oracle_shim_env_sample(S.h, n, xi1.ctypes.data_as(fp), xi2.ctypes.data_as(fp), d.ctypes.data_as(fp),
                                             pdf.ctypes.data_as(fp), le.ctypes.data_as(fp))
        le2, pdf2 = np.empty((n, 3), np.float32), np.empty(n, np.float32)
        R._lib().uivr_oracle_shim_env_eval(S.h, n, d.ctypes.data_as(fp), le2.ctypes.data_as(fp), pdf2.ctypes.data_as(fp))
    # away from patch borders (where the round trip may land in the neighbouring patch) both agree
    same = np.abs(pdf2 - pdf) <= 1e-4 * pdf
    assert same.mean() > 0.99
    assert np.max(np.abs(le2[same] - le[same]) / (1e-3 + le[same])) < 2e-3
    # the estimator of the emitted power integral is unbiased: E[Le / pdf] == integral of Le over the sphere
    est = (le.astype(np.float64) / pdf[:, None]).mean(axis=0)
    h, w = img.shape[:2]
    verts = np.concatenate([img, img[:, :1]], axis=1).astype(np.float64) * 0.7
    theta = np.arange(h) * np.pi / (h - 1)
    # quadrature of the bilinear interpolant x sin(theta): fine sub-sampling of every patch
    k = 8
    t = (np.arange(k) + 0.5) / k
    tot = np.zeros(3)
    for a in t:
        for b in t:
            val = ((1 - a) * (1 - b))[..., None] * verts[:-1, :-1] + (a * (1 - b)) * verts[:-1, 1:] + \
                  ((1 - a) * b) * verts[1:, :-1] + (a * b) * verts[1:, 1:]
            th = (np.arange(h - 1) + b) * np.pi / (h - 1)
            tot += (val * np.sin(th)[:, None, None]).sum(axis=(0, 1))
    tot *= (2 * np.pi / w) * (np.pi / (h - 1)) / (k * k)
    assert np.max(np.abs(est - tot) / tot) < 0.03


def test_envmap_nee_is_unbiased(uivr, oracle):
    """With use_nee=False the envmap only enters through Emitter::eval on escape (no sampling, no
    pdf); with NEE the sampled + MIS-weighted estimator must converge to the same image."""
    sig, alb, vol, _ = _envmap_scene(uivr, w=6, h=5)
    desc = vol.as_dict()
    spp = 6000
    a, _, _ = oracle.render_forward(desc, dict(max_depth=6, use_nee=True), sig, alb, 1, spp)
    b, _, _ = oracle.render_forward(desc, dict(max_depth=6, use_nee=False), sig, alb, 2, spp)
    assert abs(a.mean() - b.mean()) < 0.02 * b.mean()
    assert np.max(np.abs(a.mean(axis=2) - b.mean(axis=2))) < 0.12 * b.mean()
