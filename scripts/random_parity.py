"""Randomized parity sweep beyond the 16 cases of tests/test_gpu_parity.py::test_randomized_parity_sweep: seeded random
scenes (anisotropic grid / box, any film, camera, supergrid factor, flag set, emitter), CUDA slot-pool kernels
against the oracle: per-sample radiance and event counters bit-exact (forward and backward), gradients < 1e-3.

    python scripts/random_parity.py first=16 count=300
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import uivr_b200 as uivr  # noqa: E402
from oracle import oracle  # noqa: E402
from helpers import loss_grad, random_case, rel_linf  # noqa: E402
import test_gpu_parity as T  # noqa: E402


def main(first=16, count=300, variant=3):
    oracle.build()
    oracle.set_nee_log_capacity(T.NEE_LOG_CAPACITY)   # the counters model the kernel's collision log (tests/conftest.py)
    dev = torch.device("cuda:0")
    worst_s = worst_a = 0.0
    bad = 0
    for case in range(first, first + count):
        c = random_case(uivr, case)
        vol, sig, alb, props, spp = c["vol"], c["sig"], c["alb"], c["props"], c["spp"]
        quadratic = props.get("use_drt", True) and not props.get("use_drt_subsampling", True)
        v = 1 if quadratic else variant
        desc = vol.as_dict()
        img_o, smp_o, cnt_o = oracle.render_forward(desc, props, sig, alb, c["seed"], spp, want_samples=True)
        img_g, smp_g, cnt_g = T._run_forward(uivr, vol, props, sig, alb, c["seed"], spp, dev, v)
        ok = np.array_equal(smp_g.view(np.uint32), smp_o.view(np.uint32)) and cnt_g == cnt_o
        gimg = loss_grad(img_o)
        ds_o, da_o, smp_bo, cnt_bo = oracle.render_backward(desc, props, sig, alb, gimg, c["seed_grad"], spp, want_samples=True)
        cnt_bo = T._pipeline_counters(oracle, cnt_bo, v, props)
        ds_g, da_g, smp_bg, cnt_bg = T._run_backward(uivr, vol, props, sig, alb, gimg, c["seed_grad"], spp, dev, v)
        ok = ok and np.array_equal(smp_bg.view(np.uint32), smp_bo.view(np.uint32)) and cnt_bg == cnt_bo
        es = rel_linf(ds_g, ds_o) if np.abs(ds_o).max() > 0 else 0.0
        ea = rel_linf(da_g, da_o) if np.abs(da_o).max() > 0 else 0.0
        ok = ok and es < T.GRAD_TOL and ea < T.GRAD_TOL
        worst_s, worst_a = max(worst_s, es), max(worst_a, ea)
        if not ok:
            bad += 1
            print("MISMATCH case", case, props, "grad errors", es, ea, flush=True)
    print(f"{count} random cases (seeds {first}..{first + count - 1}), kernel variant {variant} vs oracle: {bad} mismatches; "
          f"worst gradient relative L-inf {worst_s:.2e} (sigma_t) / {worst_a:.2e} (albedo)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(**{k: int(v) for k, v in (a.split("=") for a in sys.argv[1:])}))
