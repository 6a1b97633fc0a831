#!/usr/bin/env python
"""bench.py -- headline benchmark of the volpathsimple hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                    [--workload config3|dense|config5] [--scaling weak|strong] [--no-extras]

A "step" = one forward render at `seed` + one DRT backward at `seed_grad` (path-replay adjoint,
which gathers the primal radiance itself, + DRT) of ONE view, i.e. what `mi.render` + `dr.backward(loss)` execute
per view in the reference (optimize.py:345-350).  Default workload: BASELINE.json configs[2]
(the configuration the metric is quoted on): 256^3 sigma_t + albedo grids, 512x512x64 spp,
`volpathsimple-drt` flags, max_depth 64, supergrid factor 8.  At N>1 the pixels are sharded
across ranks (pixel-interleaved); `--scaling weak` (default) grows spp with N so that per-GPU work
stays fixed, `strong` keeps the total; the gradients are summed with one NCCL all-reduce inside the
timed region.  Other workloads: `dense` (config-3 shapes, a medium without empty space: the case
where HBM traffic per sample is ~10x higher) and `config5` (BASELINE.json configs[4]: 512^3,
1024x1024x128 spp in total over the N ranks).

The own arm prints ONE JSON line with `value` (inputs resident in HBM), `e2e` (host buffers, copies
inside the timed region), `roofline` (dominant kernels: the backward pipeline; algorithmic bytes
from the in-kernel event counters of the same steps; `roofline.issue` = the bound that actually
binds on this path: warp-instruction issue), `cpu_baseline` (the CPU oracle on a bounded sample,
N=1 only), `clocks`, and `extra_workloads` (dense at N=1; config-3 strong scaling at N>1; config 5
at N=8): additional, clearly labelled measurements under the same timing rules.
`--impl reference` times the CPU restatement of the reference algorithm (oracle/, all host
threads): the reference itself (Dr.Jit/Mitsuba 3) cannot be installed here.  That arm
imports nothing of the product package and maps no CUDA library.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Msamples/s (fwd+DRT bwd) on 256^3 grid @512^2x64spp"
UNIT = "Msamples/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
CSRC = os.path.join(ROOT, "unbiased-inverse-volume-rendering_b200", "csrc")

# name: (grid n, film w, h, spp, dense medium)  -- the same table as oracle/workload.py
WORKLOADS = {
    "config3": (256, 512, 512, 64, False),
    "dense": (256, 512, 512, 64, True),
    "config5": (512, 1024, 1024, 128, False),
}
DESCRIPTIONS = {
    "config3": "config3: {n}^3 sigma_t+albedo, {w}x{h}x{spp}spp, DRT fwd+bwd, single view",
    "dense": "dense: config-3 shapes ({n}^3, {w}x{h}x{spp}spp, DRT fwd+bwd) on a medium WITHOUT empty space "
             "(sigma_t grid = 0.5 + 0.5 f, scale 8: optical thickness ~12 across the box)",
    "config5": "config5: {n}^3 sigma_t+albedo, {w}x{h}x{spp}spp, DRT fwd+bwd, single view, pixel-sharded",
}


def kernel_source_sha() -> str:
    """Fingerprint of the CUDA sources the shipped library is built from (stamps profiler-derived constants)."""
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except (OSError, ValueError):
        pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(cnt: dict, hw: int) -> int:
    """32 N_sigma + 96 N_alb + 4 N_maj + 64 G_sigma + 192 G_alb + 12 HW
    (12 HW = image written by the forward / grad_image read by the backward)."""
    return (32 * cnt["sigma_taps"] + 96 * cnt["albedo_taps"] + 4 * cnt["majorant_reads"] +
            64 * cnt["sigma_scatters"] + 192 * cnt["albedo_scatters"] + 12 * hw)


def profiler_constants():
    """Per-step constants that only a profiler can give (warp instructions, lanes per instruction, DRAM bytes
    of the config-3 step), captured with ncu on the GPU box and committed under profiles/ with the fingerprint
    of the kernel sources they were measured on.  Refused (None + reason) when the sources have changed since."""
    path = os.path.join(ROOT, "profiles", "kernel_constants.json")
    try:
        with open(path) as f:
            d = json.load(f)
    except (OSError, ValueError):
        return None, "profiles/kernel_constants.json missing"
    sha = kernel_source_sha()
    if d.get("source_sha") != sha:
        return None, f"stale: captured on kernel sources {d.get('source_sha')}, this tree is {sha}"
    return d, "ok"


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = f"/tmp/uivr_clocks_{os.getpid()}.csv"

    def start(self):
        """Launch `nvidia-smi -lms 25` and wait for its first sample (start-up can take longer than a short timed
        region); mark() then sets the beginning of the window whose samples count."""
        self.offset = 0
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
            return
        t0 = time.time()
        while time.time() - t0 < 3.0 and os.path.getsize(self.path) == 0:
            time.sleep(0.01)

    def mark(self):
        """Samples written from now on lie in the timed region."""
        if self.proc is not None:
            self.offset = os.path.getsize(self.path)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.close()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            f.seek(self.offset)
            for line in f:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
                except ValueError:
                    continue
                for name, v in zip(names, c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w": statistics.median(power),
                "n_samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------
# CPU arm (oracle): nothing below imports the product package
# ----------------------------------------------------------------------------------------
def _cpu_workload(name: str):
    from oracle import oracle as O
    from oracle import workload as W
    O.build()
    n, w, h, spp, dense = WORKLOADS[name]
    sig, alb = W.synthetic_grids(n, dense=dense)
    return O, W, W.benchmark_desc(n, w, h), W.drt_props(), sig, alb, (n, w, h, spp)


def cpu_baseline(name: str = "config3", target_s: float = 14.0):
    """The CPU oracle (kind 'port': C restatement of the reference algorithm, pthreads over all host cores) on a
    bounded sample of the workload: same grids / camera / flags, spp reduced so that it takes ~10-30 s."""
    O, W, desc, props, sig, alb, (n, w, h, spp_full) = _cpu_workload(name)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    W.oracle_step(O, desc, props, sig, alb, 0, 1, cores)  # calibration (also warms the page cache)
    t1 = time.perf_counter() - t0
    spp = int(max(1, min(spp_full, round(target_s / max(t1, 1e-3)))))
    t0 = time.perf_counter()
    W.oracle_step(O, desc, props, sig, alb, 1, spp, cores)
    dt = time.perf_counter() - t0
    return {"value": w * h * spp / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{name} grids/camera/flags at {w}x{h}x{spp}spp ({w * h * spp} samples, "
                      f"fwd+bwd, {dt:.1f} s of wall time on {cores} threads)"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  Mitsuba 3 / Dr.Jit (llvm_ad_rgb)
    cannot be installed here, so this is the oracle port with all host threads, on the workload's full
    configuration whenever the (K + W) steps fit in ~5 minutes (config 3 does: ~4 s per step on 16 threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    O, W, desc, props, sig, alb, (n, w, h, spp_full) = _cpu_workload(args.workload)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    W.oracle_step(O, desc, props, sig, alb, 0, 1, cores)
    W.oracle_step(O, desc, props, sig, alb, 0, 2, cores)
    t2 = time.perf_counter() - t0
    t0 = time.perf_counter()
    W.oracle_step(O, desc, props, sig, alb, 0, 4, cores)
    per_spp = max((time.perf_counter() - t0) / 4.0, 1e-4)  # marginal cost of one spp (set-up amortised)
    spp_total = spp_full * (args.gpus if args.scaling == "weak" and args.workload != "config5" else 1)
    budget = 300.0 / max(1, args.steps + args.warmup)
    spp = int(max(1, min(spp_total, budget / per_spp)))
    for it in range(args.warmup):
        W.oracle_step(O, desc, props, sig, alb, it, spp, cores)
    t0 = time.perf_counter()
    for it in range(args.warmup, args.warmup + args.steps):
        W.oracle_step(O, desc, props, sig, alb, it, spp, cores)
    dt = time.perf_counter() - t0
    samples = w * h * spp
    value = samples * args.steps / dt / 1e6
    sample = (f"each step = {args.workload} grids/camera/flags at {w}x{h}x{spp}spp ({samples} samples"
              f"{', the full configuration' if spp == spp_total else ', bounded: full is ' + str(spp_total) + ' spp'}), "
              f"fwd+bwd, CPU oracle port on {cores} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args.workload, args.gpus, spp_total, args.scaling, {"reference_sample": sample}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(name: str, n_gpus: int, spp: int, scaling: str, extra=None):
    n, w, h, _, dense = WORKLOADS[name]
    oct_mb = (n + 1) ** 3 * 32 / 1e6
    cfg = {
        "workload": DESCRIPTIONS[name].format(n=n, w=w, h=h, spp=spp),
        "integrator": "volpathsimple-drt (nee, drt, subsampling, mis), max_depth 64, majorant factor 8",
        "samples_per_step": w * h * spp,
        "parallelism": "single GPU" if n_gpus == 1 else
                       f"pixel-sharded x{n_gpus} (pixels interleaved across ranks), {scaling} scaling "
                       f"(spp {spp} in total), NCCL grad all-reduce in step",
        "l2_policy": f"inputs larger than L2 (sigma_t octets {oct_mb:.0f} MB + albedo {n ** 3 * 12 / 1e6:.0f} MB + "
                     f"gradient accumulation tiles {n ** 3 * 32 / 1e6:.0f} MB vs 126 MB L2); gradient buffers re-zeroed every step; "
                     "seeds change every step",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------------
# parity leg (N=1): CUDA path vs the oracle on the bench's own workload, in a child process
# ----------------------------------------------------------------------------------------
def parity_check(spp: int = 2):
    """BASELINE.json's metric is quoted with "grad L-inf vs ref": the CUDA path against the oracle (the
    checker, never the thing measured) on config 3's grids / camera / flags at a spp the oracle finishes in a
    second -- per-sample radiance bit for bit, image and gradient L-inf."""
    try:
        import numpy as np
        import torch
        import uivr_b200 as u
        from oracle import oracle as O
        O.build()
        n, w, h, _, _ = WORKLOADS["config3"]
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        torch.cuda.set_device(dev)
        sig_h, alb_h = u.synthetic_grids(n)
        vol = u.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=8)
        scene = u.Scene(vol, device=dev.index)
        integ = u.get_int_config("volpathsimple-drt").create(max_depth=64)
        params = {"medium.sigma_t.data": sig_h.to(dev), "medium.albedo.data": alb_h.to(dev)}
        seed, seed_grad = u.tea32(0, 1234), u.tea32(1, 1234)
        desc, props = vol.as_dict(), integ.props()
        sig, alb = sig_h.numpy(), alb_h.numpy()
        ns = w * h * spp
        nthreads = os.cpu_count() or 1
        img_o, smp_o, _ = O.render_forward(desc, props, sig, alb, seed, spp, want_samples=True, nthreads=nthreads)
        gimg = (2.0 * (img_o.astype(np.float64) - 0.5) / img_o.size).astype(np.float32)
        ds_o, da_o, smp_bo, _ = O.render_backward(desc, props, sig, alb, gimg, seed_grad, spp, want_samples=True,
                                                  nthreads=nthreads)
        smp = torch.zeros((ns, 3), device=dev)
        img = integ.render(scene, params, seed=seed, spp=spp, sample_out=smp)
        smp_b = torch.zeros((ns, 3), device=dev)
        ds, da = integ.render_backward(scene, params, torch.from_numpy(gimg).to(dev), seed=seed_grad, spp=spp,
                                       sample_out=smp_b)
        torch.cuda.synchronize()

        def rel(a, b):
            return float(np.max(np.abs(a.astype(np.float64) - b)) / max(float(np.max(np.abs(b))), 1e-30))

        return {
            "checked_against": "oracle (CPU restatement of the reference algorithm, pinned by refshim vectors)",
            "sample": f"config3 grids/camera/flags at {w}x{h}x{spp}spp ({ns} samples), matched seeds",
            "per_sample_radiance_bit_exact": bool(
                np.array_equal(smp.cpu().numpy().view(np.uint32), smp_o.view(np.uint32))
                and np.array_equal(smp_b.cpu().numpy().view(np.uint32), smp_bo.view(np.uint32))),
            "image_linf": float(np.max(np.abs(img.cpu().numpy() - img_o))),
            "grad_sigma_t_linf_rel": rel(ds.cpu().numpy(), ds_o),
            "grad_albedo_linf_rel": rel(da.cpu().numpy(), da_o),
            "tolerance": 1e-3,
        }
    except Exception as e:  # noqa: BLE001 -- the measurement must be reported whatever happens here
        return {"error": f"{type(e).__name__}: {e}"}


def parity_in_child(timeout_s: int = 240) -> dict:
    """Runs the parity leg in a child process, so that nothing it does (a crash included) can reach the
    process that holds the measurement."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--parity-only"], capture_output=True,
                           text=True, timeout=timeout_s)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": f"parity process exited with {r.returncode}: {r.stderr.strip()[-300:]}"}
        return json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
class Job:
    """One workload on this rank: grids resident on the device, the scene, the gradient buffer."""

    def __init__(self, name, scaling, world, rank, local_rank):
        import torch
        import uivr_b200 as u
        from importlib import import_module
        self.torch, self.u = torch, u
        self.sharding = import_module("uivr_b200.sharding")
        self.name, self.world, self.rank = name, world, rank
        n, w, h, spp, dense = WORKLOADS[name]
        self.n, self.w, self.h = n, w, h
        weak = scaling == "weak" and name != "config5"
        self.scaling = "weak" if weak else "strong"
        self.spp = spp * world if weak else spp
        self.dev = torch.device("cuda", local_rank)
        self.shard = self.sharding.pixel_shard(rank, world)
        self.sig_h, self.alb_h = u.synthetic_grids(n, dense=dense)
        self.sig, self.alb = self.sig_h.to(self.dev), self.alb_h.to(self.dev)
        self.vol = u.benchmark_scene(n, w, h, scale=8.0, majorant_resolution_factor=8)
        self.scene = u.Scene(self.vol, device=local_rank)
        self.integ = u.get_int_config("volpathsimple-drt").create(max_depth=64)
        self.params = {"medium.sigma_t.data": self.sig, "medium.albedo.data": self.alb}
        self.grads = self.sharding.GradientBuffer(self.vol.res, self.dev)
        self.ctx = self.scene.ctx
        self.hw = w * h
        self.inv_numel = 1.0 / (self.hw * 3)
        self.samples = self.hw * self.spp
        self.my_hw = int(self.sharding.owned_pixel_mask(self.hw, self.shard).sum().item())

    def seeds(self, it):
        return self.u.tea32(2 * it, 1234), self.u.tea32(2 * it + 1, 1234)  # optimize.py:327-328

    def step(self, it):
        seed, seed_grad = self.seeds(it)
        img = self.integ.render(self.scene, self.params, seed=seed, spp=self.spp, shard=self.shard)
        g = (img - 0.5) * (2.0 * self.inv_numel)  # d/d image of mean((image - 0.5)^2), tests:119-120
        self.integ.render_backward(self.scene, self.params, g, seed=seed_grad, spp=self.spp, shard=self.shard,
                                   out=self.grads.views())
        self.grads.all_reduce()
        return img

    def timed(self, steps, warmup, barrier, clocks=None):
        """W untimed + K timed steps, CUDA events on the launching stream, max over ranks."""
        torch = self.torch
        import torch.distributed as dist
        if clocks is not None:
            clocks.start()   # running before the warm-up: its start-up is not part of the timed region
        for it in range(warmup):
            self.step(it)
        barrier()
        if clocks is not None:
            clocks.mark()
        l0 = self.ctx.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k_fwd, k_bwd = [], []
        barrier()
        ev0.record()
        for it in range(warmup, warmup + steps):
            self.step(it)
            k_fwd.append(self.ctx.kernel_ms(0))  # waits for the kernels of this step (results are read
            k_bwd.append(self.ctx.kernel_ms(1))  # back every step in the reference loop as well)
        ev1.record()
        barrier()
        launches = self.ctx.launch_count() - l0
        clk = clocks.stop() if clocks is not None else None
        ms_total = ev0.elapsed_time(ev1)
        if self.world > 1:
            t = torch.tensor([ms_total], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
        ms_step = ms_total / steps
        return {"ms_per_step": ms_step, "value": self.samples / (ms_step * 1e-3) / 1e6, "launches": int(launches),
                "clocks": clk, "fwd_ms": sum(k_fwd) / len(k_fwd), "bwd_ms": sum(k_bwd) / len(k_bwd)}

    def count(self, steps, warmup):
        """Event counters of the timed steps (counting template instances; untimed)."""
        u, torch = self.u, self.torch
        self.ctx.set_counting(True)
        cf = dict.fromkeys(u._native.COUNTER_NAMES, 0)
        cb = dict.fromkeys(u._native.COUNTER_NAMES, 0)
        for it in range(warmup, warmup + steps):
            seed, seed_grad = self.seeds(it)
            self.ctx.reset_counters()
            img = self.integ.render(self.scene, self.params, seed=seed, spp=self.spp, shard=self.shard)
            for k, v in self.ctx.get_counters().items():
                cf[k] += v
            g = (img - 0.5) * (2.0 * self.inv_numel)
            self.ctx.reset_counters()
            self.integ.render_backward(self.scene, self.params, g, seed=seed_grad, spp=self.spp, shard=self.shard,
                                       out=self.grads.views())
            for k, v in self.ctx.get_counters().items():
                cb[k] += v
        self.ctx.set_counting(False)
        torch.cuda.synchronize()
        self.ctx.check_watchdog()  # raises if a persistent kernel aborted (results would be invalid)
        return cf, cb

    def roofline(self, t, cf, cb, steps, n_sm):
        """HBM roofline of the backward pipeline from the algorithmic bytes of these very steps, plus the
        instruction-issue bound (profiler constants; config 3 on one GPU only)."""
        peak, peak_src = hbm_peak()
        bytes_b = algorithmic_bytes(cb, self.my_hw * steps) / steps
        bytes_f = algorithmic_bytes(cf, self.my_hw * steps) / steps
        my_samples = self.my_hw * self.spp
        achieved = bytes_b / (t["bwd_ms"] * 1e-3) / 1e9
        out = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
               "traffic": None, "peak_source": peak_src,
               "kernel": "backward pipeline, rank 0: k_pool<ADJ> adjoint replay (gathers L itself) -> k_pool<DRT> -> folds "
                         "of the gradient tiles (two slot-pool launches + two streaming folds timed as one bracket)",
               "kernel_ms": t["bwd_ms"], "algorithmic_bytes_per_launch": bytes_b,
               "bytes_per_sample": bytes_b / my_samples,
               "forward_kernel": {"kernel_ms": t["fwd_ms"], "algorithmic_bytes_per_launch": bytes_f,
                                  "achieved": bytes_f / (t["fwd_ms"] * 1e-3) / 1e9,
                                  "bytes_per_sample": bytes_f / my_samples},
               "whole_step": {"algorithmic_bytes": bytes_b + bytes_f, "ms": t["fwd_ms"] + t["bwd_ms"],
                              "achieved": (bytes_b + bytes_f) / ((t["fwd_ms"] + t["bwd_ms"]) * 1e-3) / 1e9,
                              "frac": (bytes_b + bytes_f) / ((t["fwd_ms"] + t["bwd_ms"]) * 1e-3) / 1e9 / peak},
               "kernel_share_of_step": (t["bwd_ms"] + t["fwd_ms"]) / t["ms_per_step"]}
        if self.name == "config3" and self.world == 1:
            const, why = profiler_constants()
            if const is None:
                out["traffic_note"] = why
                out["issue"] = {"unavailable": why}
            else:
                out["traffic"] = const["bwd_pipeline_dram_bytes_per_step"]
                out["traffic_source"] = f"ncu dram__bytes_read+write.sum, {const.get('captured', '?')}, kernel sources {const['source_sha']}"
                clk = (t.get("clocks") or {}).get("sm_mhz") or const.get("sm_mhz_at_capture") or 1965.0
                slots = n_sm * 4 * clk * 1e6  # one warp instruction per SM sub-partition per cycle
                wi_b, wi_f = const["bwd_pipeline_warp_inst_per_step"], const["fwd_warp_inst_per_step"]
                out["issue"] = {
                    "what": "instruction-issue bound, the one that binds on this path: warp instructions executed per "
                            "step (ncu smsp__inst_executed.sum of the same kernels, fingerprinted) / (SMs x 4 "
                            "schedulers x SM clock x kernel time)",
                    "warp_inst_per_step": {"forward": wi_f, "backward_pipeline": wi_b},
                    "issue_slots_per_s": slots, "sm_mhz": clk, "n_sm": n_sm,
                    "frac_backward_pipeline": wi_b / (slots * t["bwd_ms"] * 1e-3),
                    "frac_forward": wi_f / (slots * t["fwd_ms"] * 1e-3),
                    "frac_whole_step": (wi_b + wi_f) / (slots * (t["bwd_ms"] + t["fwd_ms"]) * 1e-3),
                    "lanes_per_instruction": const.get("lanes_per_instruction"),
                    "warp_inst_per_sample": (wi_b + wi_f) / my_samples,
                }
        return out


def run_native(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    job = Job(args.workload, args.scaling, world, rank, local_rank)
    n, w, h = job.n, job.w, job.h
    spp, S_total = job.spp, job.samples

    # ---- resident-input timing: `value` ----
    t = job.timed(args.steps, args.warmup, barrier, ClockSampler(local_rank) if rank == 0 else None)

    # ---- end-to-end through host buffers: `e2e` ----
    sig_h, alb_h, sig, alb = job.sig_h, job.alb_h, job.sig, job.alb
    scene, integ, params, grads, ctx, shard = job.scene, job.integ, job.params, job.grads, job.ctx, job.shard
    inv_numel = job.inv_numel
    h_img = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
    h_g = torch.empty_like(h_img).pin_memory()
    root = rank == 0
    # only the rank that talks to the host pins the parameters / gradients (N x 0.5 GB of pinned memory otherwise)
    h_sig = sig_h.contiguous().pin_memory() if root else None
    h_alb = alb_h.contiguous().pin_memory() if root else None
    h_ds = torch.empty_like(sig_h).pin_memory() if root else None
    h_da = torch.empty_like(alb_h).pin_memory() if root else None
    stream = int(torch.cuda.current_stream().cuda_stream)
    scene.bind(None, integ.props())

    def e2e_step(it):
        seed, seed_grad = job.seeds(it)
        if world == 1:
            # the C-ABI host entry points: H2D parameters, update_medium, render, D2H image ...
            ctx.render_forward_host(h_sig.data_ptr(), h_alb.data_ptr(), seed, spp, h_img.data_ptr(), None, stream)
            torch.sub(h_img, 0.5, out=h_g).mul_(2.0 * inv_numel)  # loss gradient on the host
            # ... H2D grad_image, backward on the staged parameters, D2H gradients
            ctx.render_backward_host(None, None, h_g.data_ptr(), seed_grad, spp, h_ds.data_ptr(), h_da.data_ptr(),
                                     None, stream)
        else:
            # rank 0 alone crosses PCIe: parameters H2D once, broadcast over NVLink; gradients reduced to
            # rank 0 over NVLink, one D2H.  (Each rank copying the full tensors over the shared host link
            # was the e2e scaling limit in round 1.)
            if root:
                sig.copy_(h_sig, non_blocking=True)
                alb.copy_(h_alb, non_blocking=True)
            dist.broadcast(sig, src=0)
            dist.broadcast(alb, src=0)
            scene.update_medium(sig, force=True)
            img = integ.render(scene, params, seed=seed, spp=spp, shard=shard)
            g = (img - 0.5) * (2.0 * inv_numel)
            integ.render_backward(scene, params, g, seed=seed_grad, spp=spp, shard=shard, out=grads.views())
            dist.reduce(img, dst=0)          # disjoint partial images -> the full image on rank 0
            dist.reduce(grads.flat, dst=0)   # instead of the all-reduce: only rank 0 hands gradients to the host
            if root:
                h_img.copy_(img, non_blocking=True)
                h_ds.copy_(grads.dsigma, non_blocking=True)
                h_da.copy_(grads.dalbedo, non_blocking=True)
            torch.cuda.synchronize()

    npar = sig_h.numel() + alb_h.numel()
    h2d = npar * 4 + (h_g.numel() * 4 if world == 1 else 0)
    d2h = h_img.numel() * 4 + npar * 4
    e2e_steps = max(1, min(args.steps, 10))
    for it in range(-2, 1):   # three untimed passes (staging buffers, copy stream, first-touch of the pinned pages)
        e2e_step(it & 0xFFFF)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for it in range(1, 1 + e2e_steps):
        e2e_step(it)
    ev1.record()
    barrier()
    e2e_ms = max(ev0.elapsed_time(ev1), 0.0)
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e2e_ms, e2e_wall_ms)  # host-side loss gradient and synchronous copies count too
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_value = S_total / (e2e_ms / e2e_steps * 1e-3) / 1e6
    # restore the device-resident medium for the accounting pass
    if world > 1:
        sig.copy_(sig_h.to(dev))
        alb.copy_(alb_h.to(dev))
    scene.update_medium(sig, force=True)

    # ---- algorithmic bytes of the timed steps ----
    cnt_f, cnt_b = job.count(args.steps, args.warmup)
    roof = job.roofline(t, cnt_f, cnt_b, args.steps, n_sm)
    my_samples = job.my_hw * spp

    # ---- additional, clearly labelled workloads (same timing rules: 3 warm-up steps, CUDA events, max over ranks) ----
    extras = {}
    if not args.no_extras and args.workload == "config3" and args.scaling == "weak":
        del job, scene, params, grads
        torch.cuda.empty_cache()
        todo = [("dense", "weak")] if world == 1 else [("config3", "strong")] + ([("config5", "strong")] if world == 8 else [])
        for name, scaling in todo:
            try:
                j = Job(name, scaling, world, rank, local_rank)
                xs = 3
                tx = j.timed(xs, 3, barrier)
                cf, cb = j.count(xs, 3)
                rx = j.roofline(tx, cf, cb, xs, n_sm)
                key = name if name != "config3" else "config3_strong_scaling"
                extras[key] = {
                    "config": workload_config(name, world, j.spp, j.scaling), "value": tx["value"], "unit": UNIT,
                    "ms_per_step": tx["ms_per_step"], "steps": xs, "warmup": 3, "scaling": j.scaling,
                    "fwd_kernel_ms": tx["fwd_ms"], "bwd_pipeline_ms": tx["bwd_ms"],
                    "roofline": {k: rx[k] for k in ("bound", "achieved", "peak", "unit", "frac", "bytes_per_sample",
                                                    "whole_step")},
                    "events_per_sample": {"forward": {k: v / (j.my_hw * j.spp * xs) for k, v in cf.items()},
                                          "backward": {k: v / (j.my_hw * j.spp * xs) for k, v in cb.items()}},
                    "collective": None if world == 1 else
                                  f"one ncclAllReduce(sum, fp32) of {4 * j.n ** 3 * 4 / 1e6:.0f} MB per step, not overlapped",
                }
                del j
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001 -- an extra must never take the headline line down
                extras[name] = {"error": f"{type(e).__name__}: {e}"}
            barrier()

    if world > 1:
        dist.barrier()
    if rank == 0:
        cpu = cpu_baseline(args.workload) if (world == 1 and not args.no_cpu_baseline) else None
        line = {
            "metric": METRIC, "value": t["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t["ms_per_step"], "higher_is_better": True,
            "scaling": "weak" if (args.scaling == "weak" and args.workload != "config5") else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, world, spp, args.scaling),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "api": "uivr_render_forward_host + uivr_render_backward_host (pinned host buffers)" if world == 1
                           else "rank 0: pinned H2D + ncclBroadcast of the parameters; every rank: integrator.render / "
                                "render_backward on its pixels; ncclReduce of image + gradients to rank 0 + D2H "
                                "(h2d/d2h bytes are rank 0's; the other ranks exchange over NVLink only)"},
            "gpu_launches": t["launches"],
            "clocks": t["clocks"],
            "roofline": roof,
            "collective": None if world == 1 else
                          f"one ncclAllReduce(sum, fp32) of {4 * n ** 3 * 4 / 1e6:.0f} MB per step, after the DRT launch (not overlapped)",
            "events_per_sample": {"forward": {k: v / (my_samples * args.steps) for k, v in cnt_f.items()},
                                  "backward": {k: v / (my_samples * args.steps) for k, v in cnt_b.items()}},
        }
        if extras:
            line["extra_workloads"] = extras
        if cpu is not None:
            line["cpu_baseline"] = cpu
            line["parity"] = parity_in_child()
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line: native libraries that print there (NCCL's version banner under
    NCCL_DEBUG=VERSION) are sent to stderr for the rest of the process; emit() writes to the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="config3")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak")
    ap.add_argument("--parity-only", action="store_true", help="internal: run the parity leg and print its JSON object")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the additional workloads (extra_workloads)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        print(f"note: warmup {args.warmup} < 3 breaks the timing rules; use only for profiling runs", file=sys.stderr)
    if args.parity_only:
        _claim_stdout()
        emit(parity_check())
        return 0
    if args.impl == "reference":
        return run_reference(args)
    if "RANK" in os.environ or args.gpus <= 1:
        _claim_stdout()
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch ourselves under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
