"""GPU-vs-GPU stress of the slot-pool kernels (variant 3) against the one-sample-per-lane kernels
(variant 1) at sizes where every pool slot is recycled many times (the oracle would take too
long here; variant 1 itself is pinned to the oracle by tests/test_gpu_parity.py).

    python scripts/stress_pool.py [n=64 w=256 h=256 spp=16 factor=8]      # drt, basic and no-MIS flag sets"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uivr_b200 as u


def run(variant, n, w, h, spp, factor, combo="volpathsimple-drt", max_depth=64):
    dev = torch.device("cuda:0")
    sig, alb = u.synthetic_grids(n)
    sig, alb = sig.to(dev), alb.to(dev)
    vol = u.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=factor)
    scene = u.Scene(vol, 0)
    scene.ctx.set_variant(variant)
    integ = u.get_int_config(combo).create(max_depth=max_depth)
    params = {"m.sigma_t.data": sig, "m.albedo.data": alb}
    S = w * h * spp
    smp_f = torch.zeros((S, 3), device=dev)
    smp_b = torch.zeros((S, 3), device=dev)
    t0 = time.perf_counter()
    img = integ.render(scene, params, seed=1234, spp=spp, sample_out=smp_f)
    torch.cuda.synchronize()
    scene.ctx.check_watchdog()
    t1 = time.perf_counter()
    g = 2 * (img - 0.5) / img.numel()
    ds, da = integ.render_backward(scene, params, g, seed=u.tea32(1234, 1), spp=spp, sample_out=smp_b)
    torch.cuda.synchronize()
    scene.ctx.check_watchdog()
    t2 = time.perf_counter()
    print(f"variant {variant} n={n} {w}x{h}x{spp}: fwd {1e3 * (t1 - t0):.1f} ms bwd {1e3 * (t2 - t1):.1f} ms", flush=True)
    return img, smp_f, smp_b, ds, da


def main(n=64, w=256, h=256, spp=16, factor=8):
    rc = 0
    for combo, depth in (("volpathsimple-drt", 64), ("volpathsimple-basic", 64), ("volpathsimple-drt", 3)):
        print(f"== {combo}, max_depth {depth}")
        rc |= compare(n, w, h, spp, factor, combo, depth)
    return rc


def compare(n, w, h, spp, factor, combo, depth):
    ref = run(1, n, w, h, spp, factor, combo, depth)
    new = run(3, n, w, h, spp, factor, combo, depth)
    ok = True
    for name, a, b in zip(("image", "samples_fwd", "samples_bwd"), ref[:3], new[:3]):
        same = torch.equal(a.view(torch.int32), b.view(torch.int32)) if name != "image" else float((a - b).abs().max()) < 1e-5
        print(name, "identical" if same else f"DIFFERENT (max abs {float((a - b).abs().max()):.3e})")
        ok &= same
    for name, a, b in zip(("dsigma", "dalbedo"), ref[3:], new[3:]):
        err = float((a - b).abs().max()) / max(float(a.abs().max()), 1e-30)
        print(name, f"relative Linf {err:.3e}")
        ok &= err < 1e-4
    print("STRESS", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(**{k: int(v) for k, v in (a.split("=") for a in sys.argv[1:])}))
