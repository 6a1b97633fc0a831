"""Where does a persistent launch spend its time outside the steady state?  Needs a MEASUREMENT build of the library
(UIVR_LIB=...: csrc patched to stamp %globaltimer into the watchdog buffer: CTA start, the moment a CTA sees the
global queue exhausted, warp exit; min / max over the grid, per kernel kind).  See profiles/r02_history.md section 8b.

    UIVR_LIB=$PWD/sweep/libuivr_ts.so python scripts/ts_probe.py [spp=64 depth=64]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import uivr_b200 as u  # noqa: E402


def stamps(scene):
    buf = (C.c_uint32 * 64)()
    scene.ctx._L.uivr_check_watchdog(scene.ctx._h, buf, None)
    w = list(buf)
    out = {}
    for ki, name in enumerate(("fwd", "adj", "drt")):
        v = [w[32 + 10 * ki + 2 * f] | (w[33 + 10 * ki + 2 * f] << 32) for f in range(5)]
        if v[3] == 0:
            continue
        inv = lambda x: (~x) & 0xFFFFFFFFFFFFFFFF
        start, exh_max, exh_min, exit_max, exit_min = inv(v[0]), v[1], inv(v[2]), v[3], inv(v[4])
        out[name] = dict(total_ms=(exit_max - start) / 1e6, first_exhausted_ms=(exh_min - start) / 1e6,
                         last_exhausted_ms=(exh_max - start) / 1e6, first_cta_done_ms=(exit_min - start) / 1e6,
                         tail_after_queue_empty_ms=(exit_max - exh_min) / 1e6, exit_spread_ms=(exit_max - exit_min) / 1e6)
    return out


def main(n=256, w=512, h=512, spp=64, depth=64, factor=8):
    dev = torch.device("cuda:0")
    sig, alb = u.synthetic_grids(n)
    params = {"m.sigma_t.data": sig.to(dev), "m.albedo.data": alb.to(dev)}
    vol = u.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=factor)
    integ = u.get_int_config("volpathsimple-drt").create(max_depth=depth)
    warm = u.Scene(vol, 0)   # warm-up on another context (clocks, caches, allocations of this process)
    for _ in range(2):
        img = integ.render(warm, params, seed=1, spp=spp)
        integ.render_backward(warm, params, 2 * (img - 0.5) / img.numel(), seed=2, spp=spp)
    torch.cuda.synchronize()
    scene = u.Scene(vol, 0)   # fresh context: its debug buffer is zero, every stamp below belongs to ONE launch per kind
    img = integ.render(scene, params, seed=1234, spp=spp)
    integ.render_backward(scene, params, 2 * (img - 0.5) / img.numel(), seed=u.tea32(1234, 1), spp=spp)
    torch.cuda.synchronize()
    print(f"spp {spp} max_depth {depth}: kernel_ms fwd {scene.ctx.kernel_ms(0):.2f} bwd {scene.ctx.kernel_ms(1):.2f}")
    for k, v in stamps(scene).items():
        print(k, {a: round(b, 3) for a, b in v.items()})


if __name__ == "__main__":
    main(**{k: int(v) for k, v in (a.split("=") for a in sys.argv[1:])})
