#!/usr/bin/env python
"""bench.py -- spectral points/sec in rt_run (O2-A band, vector I/Q/U) on N B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
    python bench.py --impl reference ...                      (CPU arm: the oracle port of the
                                                               reference's CPU path, all host cores)

A "step" is one rt_run over one batch of synthetic spectral points: every Fourier
moment, every layer, surface, post-processing.  The workload is BASELINE.json
configs[1]: default_parameters() geometry/profile, Stokes_IQU, 10 streams, 34
layers, Rayleigh + one aerosol + Lambertian surface, ~10k spectral points per GPU.
Weak scaling: every rank gets its own nSpec-point shard of one nSpec*N band; the
layer control flow (ndoubl / scatter / interface) is reduced over the FULL band on
the host, each rank runs its shard with no data-path collective, and one NCCL
all-gather of R_SFI/T_SFI closes the step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "vsmartmom.jl_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np

METRIC = "spectral points/sec in rt_run (O2-A band, vector I/Q/U)"
UNIT = "spectral_points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nspec", type=int, default=10000, help="spectral points per GPU")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-aerosol", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-sample-ms", type=int, default=1000, help="nvidia-smi polling period during the timed region")
    ap.add_argument("--no-gj-compare", action="store_true",
                    help="skip the extra Gauss-Jordan-only steps (used for ncu launch lists)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


def workload_config(args, world):
    return {"workload": "configs[1]: default_parameters() O2-A band, Stokes_IQU, nstreams=10 (N=45), 34 layers, "
                        "Rayleigh+1 aerosol (HG g=0.7, 20 Greek terms -> 20 Fourier moments), Lambertian 0.15",
            "nSpec_per_gpu": args.nspec, "nSpec_total": args.nspec * world, "N": 45, "Nz": 34,
            "n_moments": 3 if args.no_aerosol else 20, "parallelism": f"nSpec-shard x{world}",
            "l2": "working set (added-layer scratch, GBs per moment) >> 126 MB L2; a 256 MiB buffer is also "
                  "rewritten between timed iterations"}


# ----------------------------------------------------------------------------
# CPU arm / cpu_baseline: the oracle port on all host cores (the only place bench.py
# touches oracle/)
# ----------------------------------------------------------------------------
def cpu_points_per_sec(model, seconds, steps=1, warmup=0, log=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vsmom_oracle as orc
    from helpers import to_oracle
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    om = to_oracle(model)
    prep = orc.prepare(om)
    pool = orc.ProcessPoolBaseline(om, cores, prep)
    chunk = 8
    rng = np.random.default_rng(0)
    pick = lambda n: np.sort(rng.choice(model.nSpec, size=min(n, model.nSpec), replace=False))
    t0 = time.perf_counter()
    pool.run(pick(cores * chunk), chunk)                       # probe (also warms the workers)
    rate = cores * chunk / (time.perf_counter() - t0)
    n = int(max(cores * chunk, min(model.nSpec, round(rate * seconds / (cores * chunk)) * cores * chunk)))
    for _ in range(warmup):
        pool.run(pick(n), chunk)
    times = []
    for _ in range(steps):
        idx = pick(n)
        t0 = time.perf_counter()
        pool.run(idx, chunk)
        times.append(time.perf_counter() - t0)
    pool.close()
    total = sum(times)
    return {"value": n * steps / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} of {model.nSpec} spectral points per step (random subset of the same band, all "
                      f"{model.m_max + 1} moments, {model.Nz} layers), {steps} step(s), {total:.1f} s; NumPy/OpenBLAS "
                      f"oracle, one forked worker per core, BLAS threads=1, dtype {np.dtype(model.float_type).name}",
            "ms_per_step": 1e3 * total / steps, "points_per_step": n}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import vsmom_b200 as vb
    FT = np.float64 if args.dtype == "f64" else np.float32
    model = vb.synthetic_o2a_model(nSpec=args.nspec, aerosol=not args.no_aerosol, float_type=FT)
    r = cpu_points_per_sec(model, args.cpu_seconds, steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference arm = CPU port of the reference's rt_run hot path (the reference is pure Julia and "
                    "cannot run here); each step is a bounded sample of the workload"}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=1000):
        self.index, self.proc, self.lines, self.period_ms = index, None, [], int(period_ms)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms), "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# ----------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------
def alg_flops(N, nd_sum, n_scat, n_adds11, n_moments, surface_adds=1):
    """Algorithmic flops per spectral point: doubling (12N^3+8N^2)/iteration,
    _11 interaction (24N^3+12N^2), elemental ~40 N^2 per layer."""
    layer = (12 * N ** 3 + 8 * N ** 2) * nd_sum + 40 * N ** 2 * n_scat
    add = (24 * N ** 3 + 12 * N ** 2) * n_adds11
    return n_moments * layer, n_moments * add + surface_adds * (24 * N ** 3 + 12 * N ** 2)


def run_b200(args):
    import torch
    import torch.distributed as dist
    import vsmom_b200 as vb
    from vsmom_b200 import _lib, rt_run as rr

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    FT = np.float64 if args.dtype == "f64" else np.float32
    tdt = torch.float64 if FT == np.float64 else torch.float32

    # one band of nspec*world points; control flow from the FULL band, then shard
    full = vb.synthetic_o2a_model(nSpec=args.nspec * world, aerosol=not args.no_aerosol, float_type=FT)
    prep = rr.prepare(full, dev)                       # Z tables + full-band layer control
    lo, hi = rank * args.nspec, (rank + 1) * args.nspec
    nS, n, nV = args.nspec, full.pol_n, len(full.vza)
    tr_h = torch.from_numpy(np.ascontiguousarray(full.tau_rayl[lo:hi].T)).pin_memory()
    ta_h = torch.from_numpy(np.ascontiguousarray(full.tau_abs[lo:hi].T)).pin_memory()
    f0 = np.zeros((nS, n), FT); f0[:, 0] = 1
    f0_h = torch.from_numpy(f0).pin_memory()
    R_h = torch.empty((nS * world, n, nV), dtype=tdt).pin_memory()
    T_h = torch.empty((nS * world, n, nV), dtype=tdt).pin_memory()

    ctx = _lib.Context(local, torch.cuda.current_stream().cuda_stream)
    tr_d, ta_d, f0_d = tr_h.to(dev), ta_h.to(dev), f0_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    Rg = torch.empty((world * nS, n, nV), dtype=tdt, device=dev) if world > 1 else None
    Tg = torch.empty_like(Rg) if world > 1 else None

    done = torch.cuda.Event(blocking=True)

    def step_device():
        R, T, _, _, _ = rr.rt_column_device(ctx, prep, tr_d, ta_d, f0_d, want_hdr=False)
        if world > 1:                                  # the single gather of R/T at the end of rt_run
            dist.all_gather_into_tensor(Rg, R)
            dist.all_gather_into_tensor(Tg, T)
            return Rg, Tg
        return R, T

    def step_e2e():
        a = tr_h.to(dev, non_blocking=True); b = ta_h.to(dev, non_blocking=True); c = f0_h.to(dev, non_blocking=True)
        R, T, _, _, _ = rr.rt_column_device(ctx, prep, a, b, c, want_hdr=False)
        if world > 1:
            dist.all_gather_into_tensor(Rg, R)
            dist.all_gather_into_tensor(Tg, T)
            R, T = Rg, Tg
        R_h.copy_(R, non_blocking=True); T_h.copy_(T, non_blocking=True)
        done.record()
        done.synchronize()          # blocking wait (the host thread sleeps instead of spinning)

    def barrier():
        # sleep on a blocking event first: a host thread that spins for seconds in
        # cudaDeviceSynchronize eats the container's CPU quota and gets descheduled afterwards
        done.record()
        done.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = ctx.kernel_launches
    # (each NVML poll stalls the GPU's work submission for ~10 ms: 200 ms polling cost 3-6 % of the step)
    sampler = ClockSampler(local, args.clock_sample_ms)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        flush.zero_()
        step_device()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = ctx.kernel_launches - launches0
    t_total, t_layer, t_add = ctx.last_column_ms()
    n_layer_launches, n_add_launches = ctx.last_column_launches()

    # e2e through the public host-buffer path
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())

    # the same step with the adaptive inverse switched off (Gauss-Jordan elimination for every
    # (I - A)^-1, the reference's getrf/getri route): reported beside the headline, not as it
    gj = None
    if world == 1 and FT == np.float64 and not args.no_gj_compare:
        ctx.set_inverse_mode(0)
        step_device()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(2):
            flush.zero_()
            step_device()
        g1.record()
        barrier()
        gms = g0.elapsed_time(g1) / 2
        gt = ctx.last_column_ms()
        gj = {"value": nS / (gms * 1e-3), "ms_per_step": gms, "layer_ms": gt[1], "adding_ms": gt[2]}
        ctx.set_inverse_mode(1)

    if rank == 0:
        ctl = prep.ctl
        N = full.quad_points.qp_muN.size
        nM = full.m_max + 1
        n11 = int(np.sum(np.asarray(ctl.iface)[1:] == 3))
        f_layer, f_add = alg_flops(N, int(ctl.ndoubl.sum()), int(ctl.scatter.sum()), n11, nM)
        esz = 8 if FT == np.float64 else 4
        peak_tf = ctx.measure_fma_peak(0 if FT == np.float64 else 1)
        ach = f_layer * nS / n_layer_launches / (t_layer / n_layer_launches * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None
        try:   # dram bytes of the dominant kernel from the committed ncu --set full capture, scaled to this launch
            tj = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
            traffic = tj["bytes_per_point_moment"] * nS * nM / n_layer_launches
        except Exception:
            pass
        value = nS * world * args.steps / (ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": nS * world * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(tr_h.numel() + ta_h.numel() + f0_h.numel()) * esz,
                    "d2h_bytes_per_step": int(R_h.numel() + T_h.numel()) * esz,
                    "path": "pinned host tau_rayl/tau_abs/F0 -> H2D -> vsm_rt_column -> (all-gather) -> D2H R_SFI/T_SFI"},
            "roofline": {
                "bound": "tensor", "kernel": "k_layer (elemental + doubling, fused)",
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf if peak_tf else None,
                "traffic": traffic,
                "peak_source": f"fp{esz * 8} FMA issue-rate microbenchmark measured live on this GPU "
                               "(vsm_measure_fma_peak; fp64 DMMA measures the same, profiles/microbench); "
                               "MEASURED_PEAKS.json holds only HBM and bf16-tensor figures. "
                               "The kernel is a dense-contraction chain on the fp64 tensor/FMA datapath, not HBM-bound.",
                "flop_accounting": "algorithmic = the reference algorithm (every (I-A)^-1 counted as one 2N^3 "
                                   "inverse). With inverse_mode=adaptive the kernels execute 0 / 4N^3 / 8N^3 flops "
                                   "for it instead when ||A||_F allows a Neumann sum of order 2 / 4 / 8 "
                                   "(remainder < 1e-20), and 21% more on every product for the 45->48 tile padding.",
                "alg_flops_per_point": f_layer + f_add, "alg_flops_per_launch": f_layer * nS / n_layer_launches,
                "launch_ms": t_layer / n_layer_launches, "launches_per_step": n_layer_launches,
                "kernel_share_of_step": t_layer / t_total,
                "hbm_frac_of_measured": (nS * (full.Nz * 2 + nV * n * 2) * esz / (t_total * 1e-3) / 1e9) / peaks["hbm_gbs"]
                if peaks.get("hbm_gbs") else None,
                "adding_kernel": {"achieved": f_add * nS / (t_add * 1e-3) / 1e12, "ms_per_step": t_add},
            },
        }
        line["config"]["inverse_mode"] = "adaptive (Neumann order 2/3/4/8 when ||A||_F^K < 1e-20, else blocked Gauss-Jordan)"
        if gj is not None:
            gj["roofline_frac"] = (f_layer * nS / (gj.pop("layer_ms") * 1e-3) / 1e12) / peak_tf if peak_tf else None
            gj["adding_ms"] = gj["adding_ms"]
            line["gauss_jordan_only"] = gj
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_points_per_sec(full, args.cpu_seconds)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
