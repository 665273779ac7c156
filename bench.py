#!/usr/bin/env python
"""bench.py -- rhs! DOF-updates/s (1/PID) of the TreeMesh DGSEM hot path, BASELINE.json config 5:
3D compressible Euler, entropy-conserving flux differencing (flux_ranocha volume + surface), polydeg 3,
TreeMesh refinement level 7 (2 097 152 elements, 134 217 728 DOF per field), periodic, weak-blast-wave IC
(reference examples/euler_ec_3d.jl:9-21 at level 7; method: reference benchmark/euler_ec_3d.jl:56-72).

A "step" is ONE rhs!(du, u, semi, t) over the whole mesh. `value` = ndofs_global * steps / device time with u
resident in HBM; `e2e` = the same through the host-vector entry point (trixib200_rhs_host: pinned host u ->
device, rhs!, du -> pinned host) each step. N > 1: the SAME level-7 mesh is partitioned along the Morton curve
over the ranks ("scaling": "strong"), halo-face traces go over NCCL send/recv inside rhs!.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--level L]

`--impl reference` times the CPU restatement of Trixi.jl's rhs! (oracle/, C++/OpenMP, all host threads) on a
bounded sample of the same workload: Julia/Trixi.jl cannot run in this image, so the oracle port is
the reference arm ("kind": "port").
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_DOF = 81.0   # 5 vars x 8 B read of u + 5 x 8 B write of du + ~1 B connectivity/Jacobian
FLOP_PER_DOF = 420.0   # algorithmic FP64 flop per DOF-update, SURVEY.md section 8(d)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--level", type=int, default=int(os.environ.get("TRIXIB200_BENCH_LEVEL", "7")))
    ap.add_argument("--e2e-steps", type=int, default=0, help="0: min(steps, 8)")
    ap.add_argument("--cpu-level", type=int, default=5, help="mesh level of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # NVML missing: record that, never fake a number
            self.nv, self.err = None, repr(e)

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no NVML samples"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_sample(level, budget_s=12.0, threads=None):
    """Times the oracle's rhs! (C++/OpenMP restatement of Trixi.jl's CPU rhs!) on the level-`level` version of the
    workload with all host threads. Returns (dof_updates_per_s, cores, sample description, ms_per_rhs)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    cores = O.max_threads() if threads is None else threads
    O.set_threads(cores)
    o = O.Oracle(ndim=3, equations="euler", polydeg=3, volume_integral="flux_differencing",
                 volume_flux="flux_ranocha", surface_flux="flux_ranocha", initial_condition="weak_blast_wave",
                 gamma=1.4, coordinates_min=(-2.0,) * 3, coordinates_max=(2.0,) * 3, initial_refinement_level=level,
                 periodicity=(True,) * 3)
    u = o.compute_coefficients(0.0)
    ndofs = o.nelements * 64
    t1 = o.time_rhs(u, warm=1, reps=1)           # seconds for one rhs!
    reps = int(max(3, min(200, budget_s / max(t1, 1e-6))))
    t = o.time_rhs(u, warm=1, reps=reps) / reps  # mean seconds per rhs!
    # the same sample on ONE thread (report both), a few calls only
    O.set_threads(1)
    r1 = int(max(2, min(10, 3.0 / max(t * cores, 1e-6))))
    t_one = o.time_rhs(u, warm=1, reps=r1) / r1
    O.set_threads(cores)
    return ndofs / t, cores, (f"3D Euler EC p=3 TreeMesh level {level} ({o.nelements} elements, {ndofs} DOF/field), "
                              f"{reps} rhs! calls after 1 warm-up, {cores} OpenMP threads"), t * 1e3, reps, ndofs / t_one


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    K, W = max(args.steps, 1), args.warmup
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    cores = O.max_threads()
    O.set_threads(cores)
    level = args.cpu_level
    o = O.Oracle(ndim=3, equations="euler", polydeg=3, volume_integral="flux_differencing",
                 volume_flux="flux_ranocha", surface_flux="flux_ranocha", initial_condition="weak_blast_wave",
                 gamma=1.4, coordinates_min=(-2.0,) * 3, coordinates_max=(2.0,) * 3, initial_refinement_level=level,
                 periodicity=(True,) * 3)
    u = o.compute_coefficients(0.0)
    ndofs = o.nelements * 64
    t = o.time_rhs(u, warm=max(W, 1), reps=K) / K    # mean seconds per rhs!
    val = ndofs / t
    sample = (f"each step = one rhs! on the level-{level} sample of the workload ({o.nelements} elements, {ndofs} "
              f"DOF/field); DOF-updates/s is size-independent on the CPU once out of cache")
    line = {"impl": "reference", "metric": "rhs! DOF-updates/s (1/PID) 3D Euler EC p=3", "value": val,
            "unit": "DOF-updates/s", "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"3D compressible Euler EC flux differencing (flux_ranocha) polydeg 3 TreeMesh "
                                   f"level {args.level} periodic, weak blast wave IC", "cpu_sample_level": level},
            "cpu_baseline": {"value": val, "unit": "DOF-updates/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "DOF-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "Trixi.jl (Julia) cannot run in this image; this is the C++/OpenMP restatement of its CPU rhs!"}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------- our arm
def build_semi(level, rank, nranks, comm_id, device):
    import trixib200 as T
    eq = T.CompressibleEulerEquations3D(1.4)
    basis = T.LobattoLegendreBasisGPU(3)
    vi = T.VolumeIntegralFluxDifferencing(T.flux_ranocha)
    solver = T.DGSEMGPU(polydeg=3, surface_flux=T.flux_ranocha, volume_integral=vi, basis=basis)
    mesh = T.TreeMesh((-2.0, -2.0, -2.0), (2.0, 2.0, 2.0), initial_refinement_level=level, periodicity=True,
                      n_cells_max=10 ** 8)
    return T.SemidiscretizationHyperbolicGPU(mesh, eq, T.initial_condition_weak_blast_wave, solver,
                                             device=device, rank=rank, nranks=nranks, comm_id=comm_id,
                                             node_coordinates=False)


def run_ours(args):
    import torch
    from trixib200 import distributed as D
    import trixib200 as T
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (libtrixib200 has no CPU fallback)")
    rank, local_rank, world = D.init_process_group()
    dist = torch.distributed if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm_id = D.broadcast_comm_id() if world > 1 else None
    K, W = max(args.steps, 1), max(args.warmup, 3)

    t0 = time.time()
    semi = build_semi(args.level, rank, world, comm_id, local_rank)
    setup_s = time.time() - t0
    ndofs_global, ndofs_local = semi.ndofsglobal(), semi.ndofs()
    u = semi.compute_coefficients_gpu(0.0, on_device=True)
    du = semi.new_vector()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput
    for _ in range(W):
        T.rhs_gpu_(du, u, semi, 0.0)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = semi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        T.rhs_gpu_(du, u, semi, 0.0)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = semi.launch_count() - l0
    ms_step = ms_total / K
    value = ndofs_global / (ms_step * 1e-3)
    if not bool(torch.isfinite(du).all().item()):
        raise SystemExit("bench.py: non-finite du")

    # ---- context: one CarpenterKennedy2N54 stage, rhs! + update kernel vs the fused trixib200_rk2n_stage
    extras = {}
    try:
        u2, tmp = semi.new_vector(), semi.new_vector().zero_()
        a, b, dt = -0.4178904745, 0.3792103129999, 1e-4

        def timed(fn, reps=min(K, 10)):
            for _ in range(2):
                fn()
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(reps):
                fn()
            f1.record()
            barrier()
            return max_over_ranks(f0.elapsed_time(f1)) / reps

        def unfused():
            T.rhs_gpu_(du, u, semi, 0.0)
            semi.rk2n_update(u2, tmp, du, a, b, dt)

        extras = {"rk2n_stage_unfused_ms": timed(unfused),
                  "rk2n_stage_fused_ms": timed(lambda: semi.rk2n_stage(u2, u, tmp, 0.0, a, b, dt)),
                  "note": "one 2N Runge-Kutta stage at the same size: rhs! + update kernel vs trixib200_rk2n_stage"}
        del u2, tmp
    except Exception as ex:  # context only: never fail the bench line over it
        extras = {"error": repr(ex)}

    # ---- end to end: host vectors in, host vectors out, every step
    e2e = None
    if not args.no_e2e:
        Ke = args.e2e_steps or min(K, 8)
        nloc = semi.nunknowns()
        u_host = torch.empty(nloc, dtype=torch.float64, pin_memory=True)
        du_host = torch.empty(nloc, dtype=torch.float64, pin_memory=True)
        u_host.copy_(u)
        for _ in range(2):
            semi.rhs_host(du_host, u_host, 0.0)
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            semi.rhs_host(du_host, u_host, 0.0)     # synchronous: returns when du_host is complete
        barrier()
        sec = max_over_ranks(time.perf_counter() - t0)
        ok = bool(torch.equal(du_host, du.cpu()))
        e2e = {"value": ndofs_global * Ke / sec, "unit": "DOF-updates/s",
               "h2d_bytes_per_step": int(nloc * 8 * world), "d2h_bytes_per_step": int(nloc * 8 * world),
               "steps": Ke, "ms_per_step": sec / Ke * 1e3, "matches_device_path": ok,
               "api": "SemidiscretizationHyperbolicGPU.rhs_host -> trixib200_rhs_host (pinned host u, du)"}
        del u_host, du_host
    clocks.stop()

    if rank != 0:
        return 0
    # ---- roofline of the dominant kernel (at N=1 the rhs! IS one launch of k_line6)
    peaks, peak_src = {}, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured"
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = BYTES_PER_DOF * ndofs_local / (ms_step * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_line6_bytes_per_dof")
        traffic = traffic * ndofs_local if traffic is not None else None
    except Exception:
        pass
    fp64_peak = None
    try:
        fp64_peak = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json"))).get("fp64_tflops")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)",
                "kernel": ("k_line6<flux_ranocha, flux_ranocha, SFV=0, 4 warps, 2 CTAs/SM, 8 pairs staged>" if semi.line3d else
                           "k_warp3d<EqEuler<3>, flux_ranocha, flux_ranocha>" if semi.warp3d else "k_fused"),
                "algorithmic_bytes_per_dof": BYTES_PER_DOF,
                "fp64": {"algorithmic_flop_per_dof": FLOP_PER_DOF,
                         "achieved_tflops": FLOP_PER_DOF * ndofs_local / (ms_step * 1e-3) / 1e12,
                         "peak_tflops_measured_dfma": fp64_peak,
                         "frac": (FLOP_PER_DOF * ndofs_local / (ms_step * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None,
                         "note": "issue-bound: time = sum of the issue costs of ~1350 FP64 (2 cycles) and ~1300 other "
                                 "instructions per element pair and scheduler (profiles/r1_line6_notes.md)"}}
    line = {"metric": "rhs! DOF-updates/s (1/PID) 3D Euler EC p=3", "value": value, "unit": "DOF-updates/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"3D compressible Euler EC flux differencing (flux_ranocha volume+surface) polydeg 3 "
                                   f"TreeMesh level {args.level} periodic [-2,2]^3, weak blast wave IC "
                                   f"(BASELINE.json configs[4])",
                       "nelements": semi.nelements_global, "ndofs_per_field": ndofs_global, "nvars": 5,
                       "partition": f"morton{world}", "l2_policy": "inputs larger than L2 "
                       f"(u+du = {2 * 40 * ndofs_local / 1e9:.2f} GB per rank vs 126 MB L2), no flush",
                       "setup_s": round(setup_s, 1)},
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks.summary(),
            "extras": extras}
    if world == 1 and not args.no_cpu:
        v, cores, sample, ms, reps, v_one = cpu_sample(args.cpu_level)
        line["cpu_baseline"] = {"value": v, "unit": "DOF-updates/s", "cores": cores, "kind": "port",
                                "sample": sample, "ms_per_rhs": ms, "value_1_thread": v_one}
    print(json.dumps(line), flush=True)
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        rc = run_reference(args)
    else:
        rc = run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
    return rc


if __name__ == "__main__":
    sys.exit(main())
