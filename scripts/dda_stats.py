"""Count the supergrid DDA iterations of the free-flight walks by kind on a BASELINE configuration, with an
instrumented build of the CPU oracle (scripts/dda_stats.c).  Analysis tooling: sizes empty-space
optimisations of the CUDA walker before GPU time is spent on them.

    python scripts/dda_stats.py [n=256] [w=512] [h=512] [spp=1] [factor=8] [cap=8] [dense=0]
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import uivr_b200 as u  # noqa: E402
from oracle import oracle as O  # noqa: E402

NAMES = ["walks", "iters", "empty", "exit_cut", "oct_free", "oct_jumps", "iso_free", "iso_jumps", "first_iter_exit"]


def main(n=256, w=512, h=512, spp=1, factor=8, cap=8, dense=0):
    so = os.path.join(ROOT, "oracle", "_build", "libdda_stats.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-ffp-contract=off", "-mfma", "-pthread", "-shared",
                           "-o", so, os.path.join(ROOT, "scripts", "dda_stats.c"), "-lm"])
    L = C.CDLL(so)
    sig_t, alb_t = u.synthetic_grids(n, dense=bool(dense)) if dense else u.synthetic_grids(n)
    sig, alb = sig_t.numpy(), alb_t.numpy()
    vol = u.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=factor)
    integ = u.get_int_config("volpathsimple-drt").create(max_depth=64)
    desc, props = vol.as_dict(), integ.props()
    O._lib = None
    # route oracle.py's entry points to the instrumented library
    L0 = O.lib()
    for name in ("uivr_oracle_render_forward", "uivr_oracle_render_backward"):
        f = getattr(L, name)
        f.argtypes = getattr(L0, name).argtypes
        f.restype = C.c_int
        setattr(L0, name, f)
    res = (C.c_int32 * 3)(n, n, n)
    sg = np.ascontiguousarray(sig.reshape(-1), dtype=np.float32)
    L.dda_stats_prepare(sg.ctypes.data_as(C.POINTER(C.c_float)), res, C.c_float(desc["scale"]), int(factor), int(cap))
    out = (C.c_ulonglong * (len(NAMES) + 65))()

    def report(tag, S):
        L.dda_stats_flush_thread()
        L.dda_stats_get(out)
        d = dict(zip(NAMES, [int(v) for v in out[:len(NAMES)]]))
        it = max(d["iters"], 1)
        print(f"== {tag}: {S} samples, {d['walks'] / S:.2f} walks/sample, {d['iters'] / S:.1f} DDA iterations/sample "
              f"({d['iters'] / max(d['walks'], 1):.1f} per walk)")
        print(f"   empty-cell iterations {100 * d['empty'] / it:.1f}% | cut by the exit mask {100 * d['exit_cut'] / it:.1f}% "
              f"(walks that start on a flagged cell: {100 * d['first_iter_exit'] / max(d['walks'], 1):.1f}%)")
        print(f"   octant-cube jumps (cap {cap}): replace {100 * d['oct_free'] / it:.1f}% with {d['oct_jumps'] / S:.2f} jumps/sample "
              f"| Chebyshev jumps: {100 * d['iso_free'] / it:.1f}% with {d['iso_jumps'] / S:.2f} jumps/sample")
        return d

    S = w * h * spp
    img, _, cf = O.render_forward(desc, props, sig, alb, 1234, spp, nthreads=1)
    report("forward", S)
    L.dda_stats_prepare(sg.ctypes.data_as(C.POINTER(C.c_float)), res, C.c_float(desc["scale"]), int(factor), int(cap))
    g = (2.0 * (img.astype(np.float64) - 0.5) / img.size).astype(np.float32)
    O.render_backward(desc, props, sig, alb, g, u.tea32(1234, 1), spp, nthreads=1)
    report("backward (primal replay + adjoint + DRT)", S)


if __name__ == "__main__":
    main(**{k: int(v) for k, v in (a.split("=") for a in sys.argv[1:])})
