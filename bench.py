#!/usr/bin/env python
"""bench.py -- particle-steps/s of the iskra hot path (gather + push + boundary + deposit + MCC
+ rho all-reduce + field solve) on N B200s, with the HBM roofline of the dominant kernel and the
CPU port timed beside it.  Contract: see the task statement / DESIGN.md "Measurement".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c4] [--particles-per-gpu P]
  python bench.py --impl reference ...      # the reference's CPU algorithm (oracle port) on host cores
Under torchrun one rank per GPU; rank 0 prints ONE JSON line.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import time

import numpy as np

_STDOUT_FD = 1

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_PARTICLE_STEP = 88.0   # read x,y,vx,vy,vz,wg + write x,y,vx,vy,vz (FP64)
METRIC = "particle-steps/sec (push+gather+deposit+MCC)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c5", "c4", "walls", "seed"])
    ap.add_argument("--particles-per-gpu", type=int, default=None)
    ap.add_argument("--cells", type=int, default=None)
    ap.add_argument("--sort-interval", type=int, default=4)
    ap.add_argument("--sort-miss", type=float, default=None, help="adaptive re-group: window-miss fraction threshold (0 = fixed interval)")
    ap.add_argument("--sort-max", type=int, default=None, help="re-group a drifting species at least every this many steps")
    ap.add_argument("--sort-full", type=int, default=None, help="force a FULL sort every this many steps (0: only when the unsorted tail exceeds 1 %% of the rows)")
    ap.add_argument("--advance-path", type=int, default=0, help="0: tile directory + incremental re-group, 1: per-warp windows + radix re-group")
    ap.add_argument("--no-lean", action="store_true", help="read and write every column (88 B per particle-step) even where v_z / wg cannot change")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-particles", type=int, default=4_000_000)
    a = ap.parse_args()
    # re-ordering policy per advance path (measured sweeps: profiles/r2_policy_sweep.txt, profiles/r1c_ncu_tiled_c5_and_walls.md):
    # tile directory (c5, c4): a re-group is one out-of-place advance launch -- every 6th step at the latest, no forced full sort;
    # per-warp windows + radix re-group (surface tracker: walls; --advance-path 1; r-z on small grids): re-group when 3 % of the rows
    # miss their window, full sort every 64 steps
    legacy = a.workload in ("walls", "seed") or a.advance_path == 1
    if a.sort_miss is None:
        a.sort_miss = 0.03 if legacy else 0.02
    if a.sort_max is None:
        a.sort_max = 64 if legacy else 6
    if a.sort_full is None:
        a.sort_full = 64 if legacy else 0
    return a


def workload_defaults(a):
    if a.workload == "c5":
        return a.particles_per_gpu or 125_000_000, a.cells or 2048
    if a.workload == "seed":
        return a.particles_per_gpu or 40_000_000, a.cells or 32
    return a.particles_per_gpu or 100_000_000, a.cells or 1024


def algo_kernel(a):
    if a.advance_path == 1 or a.workload == "walls":
        return {"walls": "k_advance_tiled<TRACK> + k_advance_tracked"}.get(a.workload, "k_advance_tiled")
    return "k_advance_tile + k_advance_list" + (" (r-z)" if a.workload == "seed" else "")


def config_dict(a, ppg, cells, n_gpus, extra=None):
    names = {"c5": "C5 shard: 2D XY RF discharge + MCC (BASELINE configs[4]), %dx%d grid, %.3g particles per GPU "
                   "(1e9 over 8 GPUs), index-slice sharding, rho all-reduce, replicated field solve",
             "c4": "C4: 2D XY two-stream (BASELINE configs[3]), %dx%d grid, %.3g particles per GPU, periodic",
             "seed": "N3 (SURVEY 8f): axisymmetric r-z column, problem/13_seed.jl geometry -- %dx%d-cell-class axial grid (65x125 nodes), "
                     "%.3g particles per GPU, axial Boris pusher, plates in z, discard dim 2, no MCC",
             "walls": "N1 (SURVEY 8f): bounded RF cell with surface tracker -- %dx%d grid, %.3g particles per GPU, "
                      "2 fixed electrodes, absorbing walls, reflective block, no MCC"}
    d = {"workload": names[a.workload] % (cells, cells, ppg), "grid_cells": [cells, cells],
         "particles_per_gpu": ppg, "particles_total": ppg * n_gpus, "sort_interval": a.sort_interval,
         "sort_policy": {"miss_threshold": a.sort_miss, "max_interval": a.sort_max, "full_interval": a.sort_full,
                         "full_sort_trigger": "unsorted tail > 1 % of the rows (ionisation appends), or no tile directory"},
         "advance_path": "tile directory + incremental re-group" if a.advance_path == 0 else "per-warp windows + radix re-group (round 1)",
         "lean": (not a.no_lean) and a.workload != "seed",
         "l2": "inputs (%.1f GB of particle columns per GPU) are far larger than the 126 MB L2" % (ppg * 52 / 1e9),
         "parallelism": "particle index slices x%d, fields replicated" % n_gpus}
    if extra:
        d.update(extra)
    return d


# ------------------------------------------------------------------------------------------------
# CPU leg: the oracle port (the reference itself is pure Julia and cannot run here)
# ------------------------------------------------------------------------------------------------
def cpu_port_run(a, cells, n_particles, steps, warmup, ppc=None, force_threads=None):
    """Times the C restatement on a bounded sample of the workload: MCC + gather + push + boundary +
    deposit on n_particles rows with the workload's particles-per-cell (the grid is shrunk to keep
    it), same tables.  The field solve is left out: the reference's dense LU cannot exist at the
    workload's grid size."""
    if ppc:
        cells = max(16, int(round(math.sqrt(n_particles / ppc))))
    from iskra_b200 import datasets
    from oracle import c_oracle as CO
    from oracle import pic_oracle as O
    Lc = CO.lib()
    cores = Lc.orc_num_threads()
    rng = np.random.default_rng(0)
    n_each = n_particles // 2
    if a.workload in ("c5", "walls"):
        dh, dt = 6.7 * 0.01 / 128, 1 / (400 * 13.56e6)
        spec = [("e-", -O.qe, O.me, 30000.0, 0.0), ("He+", O.qe, 3.99 * O.mp, 300.0, 0.0)]
        bmode = (2, 1) if a.workload == "c5" else (2, 2)
    elif a.workload == "seed":
        dh, dt = 0.08 / 32, 0.075e-9
        spec = [("e-", -O.qe, O.me, 11600.0, 0.0), ("Ar+", O.qe, 3.99 * O.mp, 300.0, 0.0)]
        bmode = (0, 2)
        cells = a.cells or 32
    else:
        w = 2 * math.pi * 9e3 * math.sqrt(2e-6 * 1e24)
        dh = 5e-3 * O.c0 / w
        dt = 0.4 * dh / 1e7 / math.sqrt(2.0)
        spec = [("e-", -O.qe, O.me, 300.0, 1e7), ("He+", O.qe, 4.002602 * O.me / 5.48579903e-04, 300.0, 0.0)]
        bmode = (1, 1)
    nx = ny = cells + 1
    if a.workload == "seed":
        nx, ny = cells + 1, min(2 * cells + 1, 8192 // (cells + 1) - 1)
    cg = CO.make_grid(nx, ny, dh, dh)
    sp = []
    for name, q, m, T, drift in spec:
        s = CO.CSpecies(int(n_each * 1.1) + 64, q, m, 1.0)
        v = rng.standard_normal((3, n_each)) * O.thermal_speed(T, m)
        if drift:
            v[0, : n_each // 2] += drift
            v[0, n_each // 2:] -= drift
        if a.workload == "seed":
            s.set(rng.random(n_each) * 0.5 * (nx - 1) * dh, (0.25 + 0.5 * rng.random(n_each)) * (ny - 1) * dh, v[0], v[1], v[2])
        else:
            s.set(rng.random(n_each) * cells * dh, rng.random(n_each) * cells * dh, v[0], v[1], v[2])
        sp.append(s)
    nn = nx * ny
    E = (rng.standard_normal(3 * nn) * 10.0)
    E[2 * nn:] = 0.0
    mccs = []
    if a.workload == "c5":
        tn = 9.64e20 * np.ones(nn)
        el = datasets.helium_electron()
        io = datasets.helium_ion()
        kinds_e = [(0, 0.0), (3, 19.82), (3, 20.61), (4, 24.587)]
        mccs.append(CO.CMcc(sp[0], [(k, thr, t[:, 0], t[:, 1], sp[1] if k == 4 else None)
                                    for (k, thr), t in zip(kinds_e, el)], 0.0, 3.99 * O.mp, 300.0, tn))
        mccs.append(CO.CMcc(sp[1], [(k, 0.0, t[:, 0], t[:, 1], None) for k, t in zip((1, 0), io)],
                            0.0, 3.99 * O.mp, 300.0, tn))
    rngc = CO.make_rng(1)
    u = np.zeros(nn)
    bm = (C.c_int32 * 2)(*bmode)

    # Fixed thread policy (the same for every --gpus N): the multi-threaded variant of the port with all host threads
    # where it exists (c5 / c4: orc_advance_mt, orc_deposit_mt), one thread for the tracker / r-z variants.  The
    # caller times both policies and reports them side by side.
    use_mt = force_threads != 1 and cores > 1 and a.workload not in ("walls", "seed")
    adv = Lc.orc_advance_mt if use_mt else (Lc.orc_advance_rz if a.workload == "seed" else Lc.orc_advance)
    dep = Lc.orc_deposit_mt if use_mt else Lc.orc_deposit
    if not use_mt:
        cores = 1

    ctr = None
    if a.workload == "walls":
        # the reference's advance! with config.tracker (track! / check!, FIFO walk): single thread like the reference
        from oracle import surfaces_oracle as SO
        og = O.CartesianGrid2(np.arange(nx) * dh, np.arange(ny) * dh)
        ost = SO.create_surface_tracker(og)
        m1 = np.zeros((nx, ny), dtype=bool)
        m1[0, :] = True
        m2 = np.zeros((nx, ny), dtype=bool)
        m2[nx - 1, :] = True
        m3 = np.zeros((nx, ny), dtype=bool)
        b0, b1 = (3 * cells) // 8, (5 * cells) // 8
        m3[b0:b1 + 1, b0:b1 + 1] = True
        SO.track_surface_(ost, m1, SO.FixedPotentialElectrode(None, 0.0))
        SO.track_surface_(ost, m2, SO.FixedPotentialElectrode(None, 0.0))
        SO.track_surface_(ost, m3, SO.create_reflective_surface())
        ctr = CO.CTracker(ost, nx, ny)

    def one_step():
        cnt = sum(s.np for s in sp)
        for m in mccs:
            m.perform(cg, E, dt, rngc, want_nu=False)
        for s in sp:
            if ctr is not None:
                ctr.advance(s, cg, E, dt, bmode=bmode)
            else:
                adv(s.ref(), C.byref(cg), CO.dp(E), C.c_double(dt), bm)
        for s in sp:
            dep(C.byref(cg), s.ref(), CO.dp(u))
        return cnt

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        done += one_step()
    el_s = time.perf_counter() - t0
    return {"value": done / el_s, "unit": "particle-steps/s", "cores": int(cores), "kind": "port",
            "sample": ("C oracle (%d thread(s)): MCC + " + ("track!/check! + " if ctr is not None else "")
                       + "gather + push + boundary + deposit on %d particles, %dx%d grid, %d steps; field solve excluded "
                       "(reference dense LU impossible at this size)") % (cores, n_particles, cells, cells, steps),
            "seconds": el_s, "steps": steps}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ppg, cells = workload_defaults(a)
    ppc = ppg / float(cells * cells)
    r = cpu_port_run(a, cells, a.cpu_particles, a.steps, a.warmup, ppc=ppc)                       # all host threads
    r1 = cpu_port_run(a, cells, a.cpu_particles, max(1, a.steps // 4), 1, ppc=ppc, force_threads=1)  # the reference's own execution model
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "particle-steps/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * r["seconds"] / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(a, ppg, cells, a.gpus, {"cpu_sample_particles": a.cpu_particles}),
            "cpu_baseline": dict({k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                                 single_thread_value=r1["value"],
                                 thread_policy="all host threads (OpenMP) for `value`, identical for every --gpus N; "
                                               "single_thread_value = one thread, the reference's execution model"),
            "e2e": {"value": r["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.p = device, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], None, set()
        for ln in out.splitlines():
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: iskra_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from iskra_b200 import workloads
    from iskra_b200 import _lib as L

    ppg, cells = workload_defaults(a)
    t_build = time.perf_counter()
    wl = (workloads.build_c5(ppg, cells, n_gpus_total=8, device=local) if a.workload == "c5"
          else workloads.build_walls(ppg, cells, device=local) if a.workload == "walls"
          else workloads.build_seed(ppg, cells, device=local) if a.workload == "seed"
          else workloads.build_c4(ppg, cells, device=local))
    rt = wl.rt
    rt.use_torch_stream()
    rt.set_advance_path(a.advance_path)
    rt.set_lean(not a.no_lean)
    wl.prepare(a.sort_interval, a.sort_miss, a.sort_max, a.sort_full)
    rt.synchronize()
    build_s = time.perf_counter() - t_build

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    wl.step(a.warmup)
    rt.synchronize()
    np_before = wl.n_particles()
    rt.profile(True)
    rt.profile_read()
    for sp in wl.kinetic():
        sp.window_stats()
    def sort_stats():
        out = {}
        for sp in wl.kinetic():
            st = (C.c_int64 * 8)()
            L.check(rt.lib.iskb_species_sort_stats(sp._h, st))
            out[sp.name] = [int(v) for v in st]
        return out
    ss0 = sort_stats()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = rt.launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    wl.step(a.steps)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    l1 = rt.launch_count()
    adv_ms, adv_launches = rt.profile_read()
    ss1 = sort_stats()
    wstats = {sp.name: [int(v) for v in sp.window_stats()] for sp in wl.kinetic()}
    rt.profile(False)
    np_after = wl.n_particles()
    rt.synchronize()
    psteps_local = 0.5 * (np_before + np_after) * a.steps

    # The unsorted tail (rows born by ionisation) joins the tile segments on every re-group launch (MOVE), so no full sort
    # recurs in steady state.  A species that was not re-grouped inside the timed steps (ions: their rows hardly move)
    # still needs one re-group each time its tail reaches 1 % of the rows; it is charged for it here with the cost of
    # a FULL sort -- an upper bound, timed right here on the live state -- divided by the interval that follows from the
    # tail growth seen in the timed region.  The amortised milliseconds are ADDED to the step time `value` comes from.
    amort_ms, amort = 0.0, {}
    if a.advance_path == 0 and a.workload in ("c5", "c4"):
        for sp in wl.kinetic():
            k = sp.name
            tail0, tail1 = ss0[k][4] - ss0[k][6], ss1[k][4] - ss1[k][6]
            growth = max(0.0, (tail1 - tail0) / float(a.steps))              # rows per step (two-step-old snapshots)
            n_rows = max(1, ss1[k][4])
            interval = (0.01 * n_rows / growth) if growth > 0 else float("inf")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(rt.lib.iskb_sort_for_deposit(sp._h, None))
            e1.record()
            torch.cuda.synchronize()
            ms_full = e0.elapsed_time(e1)
            in_window = ss1[k][0] - ss0[k][0]
            moves = ss1[k][1] - ss0[k][1]
            charged = (ms_full / interval) if (in_window == 0 and moves == 0 and interval != float("inf")) else 0.0
            amort[k] = {"full_sort_ms": ms_full, "tail_growth_rows_per_step": growth if moves == 0 else None,
                        "steady_state_interval_steps": None if (interval == float("inf") or moves) else interval,
                        "full_sorts_in_timed_region": in_window, "regroup_launches_in_timed_region": moves,
                        "charged_ms_per_step": charged}
            amort_ms += charged
    ms_measured = ms
    ms = ms + amort_ms * a.steps
    t = torch.tensor([ms, psteps_local, float(l1 - l0)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_max, psteps, launches = tmax[0].item(), tsum[1].item(), tsum[2].item()
    else:
        ms_max, psteps, launches = ms, psteps_local, float(l1 - l0)
    value = psteps / (ms_max * 1e-3)

    # roofline of the dominant kernel (fused advance), this rank
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
            "kernel": algo_kernel(a), "peak_source": peak_src}
    if adv_launches > 0 and adv_ms > 0:
        per_launch_particles = psteps_local / adv_launches
        per_launch_s = adv_ms * 1e-3 / adv_launches
        ach = ALGO_BYTES_PER_PARTICLE_STEP * per_launch_particles / per_launch_s / 1e9
        roof.update({"achieved": ach, "frac": ach / peak, "avg_launch_ms": 1e3 * per_launch_s,
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PARTICLE_STEP * per_launch_particles,
                     "kernel_share_of_step": adv_ms / ms,
                     "window_stats(window_miss,deposit_outside_window,tiles_visited,-)": wstats})
    # whole-step fraction of the HBM roofline: every kernel of the step counts against it
    step_gbs = ALGO_BYTES_PER_PARTICLE_STEP * psteps_local / (ms * 1e-3) / 1e9
    roof["step_achieved"] = step_gbs
    roof["step_frac"] = step_gbs / peak
    # re-ordering work inside the timed region: full sorts (radix sort by cell) and re-grouping advance launches
    roof["reorder_in_timed_region"] = {k: {"full_sorts": ss1[k][0] - ss0[k][0], "regroup_launches": ss1[k][1] - ss0[k][1]}
                                       for k in ss1}
    roof["full_sort_amortisation"] = amort
    roof["ms_per_step_measured"] = ms_measured / a.steps
    tr = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tr):
        try:
            roof["traffic"] = json.load(open(tr)).get(a.workload if a.advance_path == 0 else a.workload + "_path1")
        except Exception:
            pass

    # end-to-end through the reference-facing API with HOST buffers
    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(a, wl, torch, dist, world)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu_all = cpu_port_run(a, cells, a.cpu_particles, 3, 1, ppc=ppg / float(cells * cells))
        cpu_one = cpu_port_run(a, cells, a.cpu_particles, 2, 1, ppc=ppg / float(cells * cells), force_threads=1)
        cpu = {k: cpu_all[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu["single_thread_value"] = cpu_one["value"]

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(a, ppg, cells, world, {"setup_seconds": build_s}),
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(a, wl, torch, dist, world):
    """Same metric through the host-facing API: species arrays start in PINNED HOST memory; the
    timed region uploads them (iskb_species_upload), runs K steps -- each with the host->device RF
    electrode value and a device->host read of the live particle counts (what the reference's
    iteration() returns every step) -- and downloads the particle state back to the host."""
    from iskra_b200 import _lib as L
    rt = wl.rt
    kin = wl.kinetic()
    try:
        bufs = []
        for s in kin:
            n = s.np
            x = torch.empty((2, s.N), dtype=torch.float64, pin_memory=True)
            v = torch.empty((3, s.N), dtype=torch.float64, pin_memory=True)
            xn, vn = x.numpy(), v.numpy()
            L.check(rt.lib.iskb_species_download(s._h, L.ptr(xn), L.ptr(vn), None, None, s.N))
            n = s.np
            bufs.append((s, xn, vn, n))
    except Exception as ex:   # pinned allocation can fail on small hosts
        return {"value": None, "unit": "particle-steps/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                "note": "pinned host buffers unavailable: %s" % ex}
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for s, xn, vn, n in bufs:
        L.check(rt.lib.iskb_species_upload(s._h, L.ptr(xn), L.ptr(vn), None, None, n, s.N))
        h2d += n * 5 * 8
    psteps = 0
    for _ in range(a.steps):
        wl.step(1)
        h2d += 8
        cnt = 0
        for s, _, _, _ in bufs:
            cnt += s._query_np()
            d2h += 16
        psteps += cnt
    for s, xn, vn, n in bufs:
        L.check(rt.lib.iskb_species_download(s._h, L.ptr(xn), L.ptr(vn), None, None, s.N))
        d2h += s.np * 5 * 8
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    t = torch.tensor([el, float(psteps)], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        el, psteps = tm[0].item(), ts[1].item()
    return {"value": psteps / el, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d / a.steps,
            "d2h_bytes_per_step": d2h / a.steps, "seconds": el,
            "note": "solve()-style call from pinned host arrays: upload x,v once, %d steps with per-step RF value "
                    "H2D and live-count D2H, download x,v once; bytes are per-rank averages over the steps" % a.steps}


def emit(line):
    """the ONE line of stdout"""
    sys.stdout.flush()
    os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    args = parse()
    # stdout carries exactly one JSON line: until it is printed, file descriptor 1 points at stderr, so whatever a library
    # writes there ("NCCL version ..." at communicator creation, compiler chatter) cannot get in front of it
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
