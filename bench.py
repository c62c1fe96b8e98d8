#!/usr/bin/env python3
"""bench.py -- DEM steps/s (and grain-updates/s) of the B200-native stepping core on BASELINE.json's headline workload.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own kernel text on the host cores

Workload (config.workload): BASELINE.json configs[1] -- 1M three-sphere clumps (data/clumps/3_clump.csv scaled to
5 mm, Mixer-demo mass/MOI), Hertz-Mindlin with history, h = 5e-6, gravity-settled in a top-open box before timing.
A "step" is one pass of the hot path: (amortised contact-list rebuild) + contact force + owner integration.

Timing: W >= 3 untimed warm-up steps, then exactly K steps bracketed by a barrier + synchronize, CUDA events on the
launching stream, max over ranks. The per-step working set (owner state 64 MB + contact stream ~0.5 GB) is larger than
the 126 MB L2, so consecutive steps do not hit a warm L2 for the streamed arrays.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "dem-engine_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def lattice_dims(n_clumps):
    """nx*ny*nz ~= n_clumps with a bed about as deep as it is wide (100^3 for 1M)."""
    n = max(1, int(round(n_clumps ** (1.0 / 3.0))))
    nx = ny = n
    nz = max(1, int(round(n_clumps / float(nx * ny))))
    return nx, ny, nz


def build_scene(n_clumps, cd_update_freq, spacing):
    from pyapi import scenes
    nx, ny, nz = lattice_dims(n_clumps)
    sc = scenes.config2_clumps(nx, ny, nz, scale=0.005, h=5e-6, cd_update_freq=cd_update_freq, seed=4150, mu=0.2,
                               Crr=0.0, spacing=spacing)
    return sc, (nx, ny, nz)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons during the timed region through NVML (a query takes ~0.1 ms, so even a
    timed region of a few tens of milliseconds is sampled many times); falls back to polling nvidia-smi."""

    def __init__(self, index=0, period_s=0.01):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.stop_flag = index, period_s, [], False
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            try:
                ids = [int(x) for x in vis.split(",") if x.strip() != ""]
                if ids and index < len(ids):
                    phys = ids[index]
            except ValueError:
                pass
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _sample_nvml(self):
        nv = self.nv
        sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        reasons = []
        for name, bit in (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                          ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                          ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)):
            if r & bit:
                reasons.append(name)
        return sm, reasons

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        parts = [x.strip() for x in out.strip().split(",")]
        self.sm_max = float(parts[1])
        reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     parts[2:6]) if v.lower().startswith("active")]
        return float(parts[0]), reasons

    def run(self):
        while not self.stop_flag:
            try:
                sm, reasons = self._sample_nvml() if self.nv else self._sample_smi()
                self.samples.append((sm, reasons, time.perf_counter()))
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self, t0=None, t1=None):
        """Median SM clock and throttle reasons over the samples taken inside [t0, t1] (the timed region); when that
        region was too short to hold three samples, over everything sampled under the same load (timed region + the
        per-kernel timing pass that follows it on the same state), and says so."""
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        window, use = "timed region", self.samples
        if t0 is not None:
            inside = [s for s in self.samples if t0 <= s[2] <= t1]
            if len(inside) >= 3:
                use = inside
            else:
                window = "timed region + per-kernel timing pass (timed region shorter than 3 samples)"
        sm = sorted(s[0] for s in use)
        reasons = sorted({r for s in use for r in s[1]})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": getattr(self, "sm_max", None), "reasons": reasons,
                "samples": len(sm), "window": window, "source": "nvml" if self.nv else "nvidia-smi"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def traffic_from_profiles():
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/), if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("k_force_ss_dram_bytes_per_launch")
    except Exception:
        return None


def cpu_reference_run(n_clumps, steps, warmup, cd_update_freq, spacing, settle_steps, budget_s=20.0, settle_budget_s=60.0):
    """Times the reference's own kernel text (oracle/_ref, host-compiled, OpenMP over all host cores) -- or the C port
    when _ref is absent -- on the workload's bed of n_clumps clumps.  The CPU cannot settle a 1M-clump bed in minutes
    (80 000 steps at ~0.1 s each), so settling is bounded by settle_budget_s of wall time and the description says how far
    it got.  Returns (steps_per_s, kind, cores, description, steps_run, seconds, n_contacts)."""
    from oracle import pyoracle
    from pyapi import scenes
    sc, dims = build_scene(n_clumps, cd_update_freq, spacing)
    f = scenes.flatten(sc)
    w = pyoracle.world_from_flat(f)
    use_ref = pyoracle.ref() is not None
    cores = os.cpu_count() or 1
    if use_ref:
        pyoracle.ref_set_threads(cores)
    else:
        cores = 1
    chunk = max(1, cd_update_freq)
    t0 = time.perf_counter()
    settled = 0
    while settled < settle_steps and time.perf_counter() - t0 < settle_budget_s:
        w.step(chunk, cd_every=cd_update_freq, use_ref=use_ref)
        settled += chunk
    if warmup:
        w.step(warmup, cd_every=cd_update_freq, use_ref=use_ref)
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        n = min(chunk, steps - done)
        w.step(n, cd_every=cd_update_freq, use_ref=use_ref)
        done += n
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, ("reference" if use_ref else "port"), cores, (
        "%d clumps (%dx%dx%d lattice, the workload's bed), %d timed steps after %d settling + %d warm-up steps on the host "
        "(%d contacts listed at the end), the reference's own force / accumulation / integration kernels on %d host "
        "threads (contact rebuild: the oracle's grid search spread over the same threads, merge of the sorted runs and history map serial)" % (f.nClumps, dims[0], dims[1], dims[2], done, settled, warmup, int(w.nContacts), cores)), done, dt, int(w.nContacts)


def run_c5(args, rank, local_rank, world):
    """BASELINE configs[4]: 5M monodisperse spheres, binning + sort only.  The spheres are pre-partitioned into x-slabs,
    one per GPU; there is no exchange in steady state, so the ranks run independently (no data-path collective)."""
    import torch
    from pyapi import demb200, dist_util, scenes
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist_util.init("nccl", local_rank)
    sc = scenes.config5_spheres(args.spheres, x_range=(rank / world, (rank + 1) / world))
    f = scenes.flatten(sc)
    eng = demb200.Engine(local_rank)
    eng.load_flat(f, contact_capacity=1024)
    eng.profile_binning(max(args.warmup, 3))
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    reps = max(args.steps // 10, 10)
    t = eng.profile_binning(reps)
    (us,) = dist_util.max_over_ranks([t["total_us"]], device="cuda")
    n_local = int(f.nSpheres)
    n_total = int(dist_util.gather_counts([n_local], device="cuda").sum()) if world > 1 else n_local
    if rank == 0:
        peak, which = measured_peak()
        algo = 96.0 * n_local  # 16 N position read + 8 N key/value write + 8 N (2p+1) sort traffic, p = 4
        line = {"metric": "sphere-keys/s, binning + sort only (5M monodisperse spheres)", "value": n_total / (us * 1e-6),
                "unit": "keys/s", "n_gpus": args.gpus, "steps": reps, "warmup": max(args.warmup, 3), "ms_per_step": us / 1000.0,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 keys (f64 position decode)",
                "data": "synthetic",
                "config": {"workload": "C5: %d spheres r=1mm uniform random at 50%% packing, binning + sort only" % n_total,
                           "spheres_per_gpu": n_local, "cells": [int(v) for v in eng.stats().n_cells],
                           "parallelism": "1 GPU" if world == 1 else "%d pre-partitioned x-slabs, no exchange" % world},
                "kernel_us": t,
                "roofline": {"bound": "hbm", "kernel": "keys + counting sort + gather", "achieved": algo / (us * 1e-6) / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": algo / (us * 1e-6) / 1e9 / peak, "traffic": None,
                             "peak_source": which, "algorithmic_bytes_per_launch": algo},
                "gpu_launches": int(eng.stats().kernel_launches)}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=40)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clumps", type=int, default=1000000)
    ap.add_argument("--settle-steps", type=int, default=80000,
                    help="untimed gravity-settling steps before warm-up (0.4 s of simulated time: the bed is at rest)")
    ap.add_argument("--cd-update-freq", type=int, default=20)
    ap.add_argument("--spacing", type=float, default=2.7, help="initial lattice spacing in units of the clump scale")
    ap.add_argument("--cpu-clumps", type=int, default=0,
                    help="clumps of the CPU runs (--impl reference and the cpu_baseline leg); 0 = the workload's full size")
    ap.add_argument("--reference-gpu", action="store_true",
                    help="also time the UNMODIFIED reference (baseline/_ref/run_ref) on this box's GPU(s): adds ~15 minutes "
                         "(its Initialize() of the 1M-clump bed takes ~7 minutes per run); without it the recorded run "
                         "profiles/reference_gpu_r02.json is quoted")
    ap.add_argument("--no-facade", action="store_true", help="skip the legs that go through the C++ deme::DEMSolver facade")
    ap.add_argument("--reference-gpu-steps", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--clock-period", type=float, default=0.01, help="seconds between NVML clock samples")
    ap.add_argument("--profile-window", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (for `ncu --profile-from-start off`)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2: the headline bed (default); c5: 5M-sphere binning + sort stress case (BASELINE configs[4])")
    ap.add_argument("--spheres", type=int, default=5000000, help="c5: total number of spheres")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    metric, unit = "DEM steps/sec for 1M 3-sphere clumps (Hertz-Mindlin with history)", "steps/s"
    workload = "C2: %d three-sphere clumps (3_clump.csv @5mm), gravity-settled bed in a top-open box, h=5e-6" % args.clumps

    # ------------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        # the reference's CPU leg of the path on this box's host cores, on the workload's own bed (full size unless
        # --cpu-clumps says otherwise); every step is one pass over all of it
        n_cpu = args.cpu_clumps if args.cpu_clumps > 0 else args.clumps
        value, kind, cores, desc, done, dt, n_cnt = cpu_reference_run(
            n_cpu, max(args.steps, 1), min(args.warmup, 5), args.cd_update_freq, args.spacing,
            settle_steps=args.settle_steps, budget_s=90.0, settle_budget_s=60.0)
        value *= n_cpu / float(args.clumps)  # (identity at full size)
        line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
                "steps": done, "warmup": min(args.warmup, 5), "ms_per_step": 1000.0 / value, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32 (f64 centre distance)", "data": "synthetic",
                "config": {"workload": workload, "lattice": list(lattice_dims(args.clumps)), "clumps_total": args.clumps,
                           "cd_update_freq": args.cd_update_freq, "force_record": False,
                           "cpu_sample_clumps": n_cpu, "contacts_listed": n_cnt},
                "grain_updates_per_s": value * args.clumps,
                "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": desc},
                "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    from pyapi import demb200, dist_util, scenes

    if args.workload == "c5":
        return run_c5(args, rank, local_rank, world)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist_util.init("nccl", local_rank)

    # strong scaling: ONE bed of args.clumps clumps, split into `world` x-slabs with a ghost-owner halo exchange every
    # step (multi-GPU). Every rank builds the same complete input; ownership follows positions.
    sc, dims = build_scene(args.clumps, args.cd_update_freq, args.spacing)
    f = scenes.flatten(sc)
    eng = demb200.Engine(local_rank)
    stream = torch.cuda.Stream(device=local_rank)
    eng.set_stream(stream.cuda_stream)
    eng.load_flat(f, contact_capacity=0 if world == 1 else int(f.nSpheres) * 6 // world + 200000)
    if world > 1:
        eng.mgpu_init(rank, world, dist_util.share_bytes(demb200.Engine.mgpu_unique_id, 128, device="cuda"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        eng.step(args.settle_steps)
        eng.step(args.warmup)
        barrier()
        launches0 = eng.stats().kernel_launches
        sampler = ClockSampler(local_rank, args.clock_period)
        sampler.start()
        t_timed0 = time.perf_counter()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.profile_window:
            torch.cuda.profiler.start()
        if world > 1:
            eng.mgpu_barrier()  # device-side rank barrier on the stream: the timed region starts with all GPUs level
        ev0.record(stream)
        eng.step_async(args.steps)
        ev1.record(stream)
        barrier()
        t_timed1 = time.perf_counter()
        if args.profile_window:
            torch.cuda.profiler.stop()
        ms = ev0.elapsed_time(ev1)
        launches = eng.stats().kernel_launches - launches0
        st = eng.stats()
        # per-kernel device times (CUDA events on the launching stream), same state, right after the timed region
        prof = eng.profile_steps(min(args.steps, 200))
        sampler.stop_flag = True

        # ---- e2e: through the public C-ABI calls with HOST buffers; every step uploads the step's host-side inputs
        # (simulation parameters + the family prescription / mask tables a co-simulating caller updates) and reads the
        # step's result metrics (max |v|, kinetic energy) back to the host.
        h2d = 56576 + 40  # family blob (masks 32896 padded to 33024 + 1024 + 88*256) + neutral elements of the reductions
        d2h = 40          # five doubles of dem_reduce_many
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            eng.set_params(eng.params)
            eng.update_families(f.familyMasks, f.familyExtraMarginSize, f.prescriptions)
            eng.step_async(1)
            eng.reduce_many((demb200.REDUCE_MAX_ABSV, demb200.REDUCE_KINETIC_ENERGY))  # synchronises
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0

    ms, e2e_ms = dist_util.max_over_ranks([ms, e2e_s * 1000.0], device="cuda")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = args.steps / (ms / 1000.0)
    n_total_clumps = f.nClumps
    mg = eng.mgpu_info()
    C_ss, C_sa = int(st.n_contacts_ss), int(st.n_contacts_sa)
    peak, which = measured_peak()
    # algorithmic bytes of the dominant kernel (k_force_ss), SURVEY.md 8(d): 41 B per candidate contact
    # (9 ids + 16 history read + 16 history write) + 7 B per sphere + (57 read + 24 accumulate) B per owner
    n_own_local = mg["n_active"] if world > 1 else f.nOwners  # owners this rank actually touches
    n_sph_local = 3 * n_own_local if world > 1 else f.nSpheres
    algo_bytes = 41.0 * C_ss + 7.0 * n_sph_local + 81.0 * n_own_local
    ach = algo_bytes / (prof["force_ss_us"] * 1e-6) / 1e9 if prof["force_ss_us"] > 0 else 0.0
    step_bytes = 41.0 * (C_ss + C_sa) + 7.0 * n_sph_local + 218.0 * n_own_local
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (f64 centre distance)", "data": "synthetic",
        "config": {"workload": workload, "lattice": list(dims), "clumps_total": n_total_clumps,
                   "spheres_per_gpu": int(n_sph_local), "contacts_ss": C_ss, "contacts_ss_touching": int(st.n_contacts_ss_touching), "contacts_sa": C_sa,
                   "cd_update_freq": args.cd_update_freq, "settle_steps": args.settle_steps,
                   "force_record": False, "l2": "per-step working set > 126 MB L2 (no flush needed)",
                   "parallelism": "1 GPU" if world == 1 else
                   "%d x-slabs, ghost-owner halo exchange per step over %s (rank 0: %d own + %d ghost owners, %d B sent per step)"
                   % (world, "NVLink peer stores + flags" if mg.get("peer_memory_exchange") else "ncclSend/ncclRecv",
                      mg["n_own"], mg["n_active"] - mg["n_own"], mg["halo_bytes_per_step"])},
        "grain_updates_per_s": value * n_total_clumps,
        "kernel_us": prof,
        "step_algorithmic_GBps": step_bytes / (ms / args.steps * 1e-3) / 1e9,
        "roofline": {"bound": "hbm", "kernel": "k_force_ss", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": traffic_from_profiles(), "peak_source": which,
                     "algorithmic_bytes_per_launch": algo_bytes},
        "e2e": {"value": args.e2e_steps / (e2e_ms / 1000.0), "unit": unit, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": args.e2e_steps},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(t_timed0, t_timed1),
    }
    if not args.no_cpu_baseline and world == 1:
        # bounded: ~20 s of timed CPU work on the full-size bed (a few hundred ms per step on the host cores)
        n_cpu = args.cpu_clumps if args.cpu_clumps > 0 else args.clumps
        try:
            rate, kind, cores, desc, done, dt, n_cnt = cpu_reference_run(n_cpu, 40, 2, args.cd_update_freq, args.spacing,
                                                                        settle_steps=args.settle_steps, budget_s=20.0,
                                                                        settle_budget_s=25.0)
            line["cpu_baseline"] = {"value": rate * n_cpu / float(args.clumps), "unit": unit, "cores": cores, "kind": kind,
                                    "sample": desc}
        except Exception as e:  # the GPU numbers above stand on their own: report the failure instead of losing the line
            line["cpu_baseline"] = {"value": None, "unit": unit, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
    if world == 1 and not args.no_facade:
        # The same bed through the reference's own C++ API: baseline/run_ref.cpp -- the driver script of the reference arm,
        # UNCHANGED -- compiled against this repository's deme::DEMSolver facade (dem-engine_b200/host/run_b200).
        #   bench: Initialize(), DoDynamicsThenSync(warm-up), then wall clock around DoDynamicsThenSync(steps * h)
        #   e2e  : every frame DoDynamics(h), a tracked clump's Pos / Vel and two inspectors read back to the host
        from tools import run_reference_gpu as rr
        import tempfile
        path = os.path.join(tempfile.gettempdir(), "dem_c2_settled_rank0.bin")
        try:
            rr.dump_settled_scene(eng, sc, f, path)
            eng.close()
            exe = os.path.join(ROOT, "dem-engine_b200", "host", "run_b200")
            fac = {}
            for mode, n in (("bench", args.steps), ("e2e", args.e2e_steps)):
                d = rr.run_reference(path, 1, n, 200, cd_update_freq=args.cd_update_freq, timeout=600, exe=exe, mode=mode)
                d.pop("stats_tail", None)
                fac[mode] = d
            line["facade"] = {"driver": "baseline/run_ref.cpp compiled against dem-engine_b200/host (run_b200)",
                              "DoDynamicsThenSync_steps_per_s": fac["bench"].get("steps_per_s"),
                              "e2e_steps_per_s": fac["e2e"].get("steps_per_s"), "runs": fac}
        except Exception as e:  # (a failed facade leg must not cost the line its measured numbers)
            line["facade"] = {"driver": "baseline/run_ref.cpp compiled against dem-engine_b200/host (run_b200)",
                              "failed": "%r" % (e,)}
        if args.reference_gpu:
            # the UNMODIFIED reference (DEMSolver(1), and DEMSolver(2) when the box has a second GPU) on the same settled
            # bed through the same driver script (baseline/_ref/run_ref); its Initialize() of 1M clumps alone takes ~7
            # minutes, which is why this leg is opt-in and the recorded run is quoted otherwise
            ref = {}
            for g in (1, 2):
                if g > torch.cuda.device_count():
                    ref["DEMSolver(%d)" % g] = {"unavailable": "box has %d GPU(s)" % torch.cuda.device_count()}
                    continue
                d = rr.run_reference(path, g, args.reference_gpu_steps, 100, timeout=1500)
                d.pop("stats_tail", None)
                ref["DEMSolver(%d)" % g] = d
            line["reference_gpu"] = ref
    if "reference_gpu" not in line:
        try:
            with open(os.path.join(ROOT, "profiles", "reference_gpu_r02.json")) as fh:
                rec = json.load(fh)
            rec["recorded"] = True
            line["reference_gpu"] = rec
        except Exception:
            line["reference_gpu"] = {"unavailable": "not measured in this run (--reference-gpu) and no recorded run under profiles/"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
