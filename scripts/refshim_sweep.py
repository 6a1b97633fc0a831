"""Wide sweep of the oracle against the reference's own files (oracle/refshim.py runs volpathsimple.py
unmodified): `python scripts/refshim_sweep.py [first=0] [count=96]` -> one line per random scene of
tests/helpers.random_case + a summary.  Needs /root/reference (build container only); evidence, not a test."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import uivr_b200 as u  # noqa: E402
from helpers import loss_grad, random_case, rel_linf  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import refshim as R  # noqa: E402


def main(first=0, count=96):
    O.build()
    worst = dict(sample=0.0, sample_bwd=0.0, dsigma=0.0, dalbedo=0.0)
    t0 = time.time()
    for case in range(first, first + count):
        c = random_case(u, case)
        vol, sig, alb, spp = c["vol"], c["sig"], c["alb"], c["spp"]
        props = dict(c["props"])
        desc = vol.as_dict()
        integ = R.make_integrator("volpathsimple-drt", max_depth=props.pop("max_depth"), **props)
        img, smp = R.render_forward(desc, integ, sig, alb, c["seed"], spp)
        gimg = loss_grad(img)
        ds, da, smp_b = R.render_backward(desc, integ, sig, alb, gimg, c["seed_grad"], spp)
        _, smp_o, _ = O.render_forward(desc, c["props"], sig, alb, c["seed"], spp, want_samples=True)
        ds_o, da_o, smp_bo, _ = O.render_backward(desc, c["props"], sig, alb, gimg, c["seed_grad"], spp, want_samples=True)
        scale = max(1.0, float(np.abs(smp).max()))
        r = dict(sample=float(np.abs(smp - smp_o).max()) / scale, sample_bwd=float(np.abs(smp_b - smp_bo).max()) / scale,
                 dsigma=rel_linf(ds_o, ds) if np.abs(ds).max() > 0 else 0.0,
                 dalbedo=rel_linf(da_o, da) if np.abs(da).max() > 0 else 0.0)
        for k, v in r.items():
            worst[k] = max(worst[k], v)
        flags = "".join(ch for ch, on in (("N", c["props"].get("use_nee")), ("D", c["props"].get("use_drt")),
                                         ("S", c["props"].get("use_drt_subsampling", True)), ("M", c["props"].get("use_drt_mis", True)),
                                         ("H", c["props"].get("hide_emitters")), ("E", vol.envmap is not None)) if on)
        print(f"case {case:3d} res {str(vol.res):14s} film {desc['width']:2d}x{desc['height']:2d} spp {spp} depth {c['props']['max_depth']:2d} "
              f"{flags:6s} sample {r['sample']:.1e} {r['sample_bwd']:.1e}  dsigma {r['dsigma']:.1e}  dalbedo {r['dalbedo']:.1e}", flush=True)
    print(f"{count} scenes in {time.time() - t0:.0f} s; worst: " + ", ".join(f"{k} {v:.1e}" for k, v in worst.items()))


if __name__ == "__main__":
    main(**{k: int(v) for k, v in (a.split("=") for a in sys.argv[1:])})
