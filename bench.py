#!/usr/bin/env python
"""bench.py -- pair-force particle-steps/s on BASELINE.json's workload, with roofline,
CPU baseline, end-to-end number and clocks (contract: task statement section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C2]

* one "step" = one full force evaluation of the workload's potentials over all particles
  (neighbour-list build excluded; multi-GPU steps include the per-step halo exchange);
* N = 1: C2 (ExpandedYukawa, N = 1,000,000, 2 types, r_cut 3.5, shift, force+energy+virial,
  fp32) -- the configuration the metric is quoted on; N > 1: the same fluid at N = n_gpus x
  1,000,000 particles (weak scaling), one contiguous Morton slice of rows per rank, halo
  positions exchanged every step with NCCL;
* `value`: whole-job particle-steps/s with inputs resident in HBM (CUDA events, max over ranks);
* `e2e`: the same metric through the public `pair` API with HOST buffers: every step copies the
  positions from pinned host memory to the device and the forces (+virial) back;
* `roofline`: algorithmic bytes / measured kernel time vs MEASURED_PEAKS.json;
* `cpu_baseline` / `--impl reference`: the CPU oracle (oracle/_ref = the reference's own evaluator
  headers under the restated HOOMD loop, else the port) on the host cores, bounded row sample.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


METRIC = "pair_force_particle_steps_per_s"
UNIT = "particle-steps/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
# committed `ncu --set full` capture of this command (profiles/); None until captured.
NCU_TRAFFIC_BYTES = {"C2": 614.97e6}  # profiles/r01_c2_ncu_full_v7_summary.csv (578.60 + 36.37 MB)
# warp instructions executed per launch of the same capture (smsp__inst_executed.sum): the kernel
# is instruction-issue bound, not HBM bound, so the issue floor is reported beside the HBM roofline
NCU_WARP_INSTRUCTIONS = {"C2": 217.95e6}
SM_COUNT, SCHEDULERS_PER_SM = 148, 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--n-per-gpu", type=int, default=0, help="override particles per GPU")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU halo transport: NVLink peer-memory push (default) or NCCL send/recv")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tune", action="store_true")
    ap.add_argument("--block", type=int, default=0, help="pin block_size (with --no-tune)")
    ap.add_argument("--tpp", type=int, default=0, help="pin threads_per_particle (with --no-tune)")
    return ap.parse_args()


DEFAULT_N = {"C1": 32000, "C2": 1000000, "C3": 4000000, "C4": 8000000, "C5": 16000000}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for n, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def timed_with_clocks(fn, device_index, probe=None):
    """Run `fn` (the timed region) with the clock sampler on; `probe`, if given, then repeats the
    same load untimed for about a second so that the sampler (one nvidia-smi call per ~0.1 s)
    sees the clocks under that load more than once -- the timed region itself lasts ~15 ms."""
    sampler = ClockSampler(device_index)
    sampler.start()
    try:
        out = fn()
        if probe is not None:
            probe()
    finally:
        sampler.stop_flag.set()
        sampler.join(timeout=2)
    return out, sampler.summary()


# -------------------------------------------------------------------------------------------------
# CPU side: the oracle on the host cores (cpu_baseline of the b200 arm; the whole reference arm)
# -------------------------------------------------------------------------------------------------
def cpu_oracle_rate(wl, target_seconds=8.0, steps=1, warmup=0, quiet=True):
    """particle-steps/s of the oracle loop on a bounded sample of rows of `wl` (fp32, full list,
    all host threads). Returns (value, info dict)."""
    from oracle import oracle

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    orc = oracle.load("best", np.float32)
    kind = "reference" if orc.kind == "ref" else "port"
    cores = orc.max_threads()
    from azplugins_b200.state import pack_pos

    pos = pack_pos(wl.position, wl.typeid, np.float32)
    spec = wl.potentials
    nt = len(wl.types)

    def tables(s):
        name = {"PerturbedLennardJones": "PerturbedLennardJones", "ExpandedYukawa": "ExpandedYukawa",
                "DPDGeneralWeight": "DPDGeneralWeight", "TwoPatchMorse": "TwoPatchMorse",
                "Colloid": "Colloid", "Hertz": "Hertz"}[s["cls"]]
        rc = np.full((nt, nt), float(s["default_r_cut"]))
        pp = {}
        for (a, b), p in s["params"].items():
            pp[(wl.types.index(a), wl.types.index(b))] = p
        for (a, b), r in s.get("r_cut", {}).items():
            i, j = wl.types.index(a), wl.types.index(b)
            rc[i, j] = rc[j, i] = r
        return name, orc.pack_table(name, nt, pp), rc

    tabs = [tables(s) for s in spec]
    rc_max = np.max([t[2] for t in tabs], axis=0)

    lists = {}

    def run(n_rows, nthreads=0):
        if n_rows not in lists:
            lists.clear()
            lists[n_rows] = orc.build_nlist(pos, wl.box.L, rc_max + 0.4, ntypes=nt, n_rows=n_rows)
        nn, nl, head = lists[n_rows]
        t0 = time.perf_counter()
        for s, (name, table, rc) in zip(spec, tabs):
            mode = s.get("kwargs", {}).get("mode", "none")
            if name == "TwoPatchMorse":
                orc.aniso_forces(table, pos, wl.orientation.astype(np.float32), nn, nl, head,
                                 wl.box.L, rc, ntypes=nt, mode=mode, virial=wl.compute_virial,
                                 N=n_rows, nthreads=nthreads)
            elif name == "DPDGeneralWeight":
                vel = np.zeros((len(pos), 4), dtype=np.float32)
                vel[:, :3] = wl.velocity
                orc.dpd_forces(table, pos, vel, wl.tag, nn, nl, head, wl.box.L, rc, wl.seed,
                               wl.timestep, wl.dt, s["kwargs"]["kT"], ntypes=nt,
                               virial=wl.compute_virial, N=n_rows, nthreads=nthreads)
            else:
                orc.pair_forces(name, table, pos, nn, nl, head, wl.box.L, rc, ntypes=nt,
                                mode=mode, virial=wl.compute_virial, N=n_rows, nthreads=nthreads)
        return time.perf_counter() - t0

    probe_rows = min(wl.N, 20000)
    t_probe = run(probe_rows)
    n_rows = int(min(wl.N, max(probe_rows, probe_rows * target_seconds / max(t_probe, 1e-6))))
    for _ in range(warmup):
        run(n_rows)
    times = [run(n_rows) for _ in range(max(1, steps))]
    t = float(np.median(times))
    # one thread on the same rows (what a single HOOMD CPU rank does); the list
    # of the sample is reused, the pass is bounded by taking the probe's rows
    t1 = run(n_rows, nthreads=1) if n_rows <= 4 * probe_rows else None
    if t1 is None:
        lists.clear()
        t1_rows = probe_rows
        t1 = run(t1_rows, nthreads=1)
    else:
        t1_rows = n_rows
    info = {"value": n_rows / t, "unit": UNIT, "cores": cores, "kind": kind,
            "value_1_thread": t1_rows / t1,
            "sample": "first %d of %d rows of %s, full neighbour list, fp32, %.2f s per pass "
                      "(HOOMD's CPU classes would use a half list: half the pair evaluations)"
                      % (n_rows, wl.N, wl.name, t)}
    return n_rows / t, info, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from azplugins_b200 import synth  # numpy-only generator

    n = args.n_per_gpu or DEFAULT_N[args.workload]
    wl = synth.CONFIGS[args.workload](N=n * max(1, args.gpus) if args.gpus > 1 else n)
    value, info, t = cpu_oracle_rate(wl, target_seconds=4.0, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "N": wl.N, "note": "CPU oracle on host cores"},
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# -------------------------------------------------------------------------------------------------
# B200 arm
# -------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import azplugins_b200 as az
    from azplugins_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    multi = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if multi:
        # NCCL's internal stream must outrank the force kernels, or its few CTAs queue behind the
        # thousands of CTAs of the interior-row kernel and the exchange serialises with the compute
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    n_per = args.n_per_gpu or DEFAULT_N[args.workload]
    wl = synth.CONFIGS[args.workload](N=n_per * world)
    n_total = wl.N
    K, W = args.steps, max(3, args.warmup)
    hbm_peak, peak_src = peaks()

    if not multi:
        state = wl.make_state(dtype=np.float32, device=dev)
        nl = az.nlist.Cell(buffer=synth.BUFFER)
        pots = wl.make_potentials(nl)
        for p in pots:
            p.attach(state)
        nl.compute(state)
        # the list is frozen for the run: the per-step displacement check (a device->host flag
        # read) belongs to the neighbour-list row, not to the force path that is timed here
        nl.freeze()
        torch.cuda.synchronize()
        tuned = []
        for p in pots:
            if not args.no_tune:
                tuned.append(p.tune_kernel_parameters(compute_virial=wl.compute_virial))
            else:
                p.kernel_parameters = (args.block, args.tpp)
                tuned.append(p.kernel_parameters + (None,))

        def step():
            for p in pots:
                p.compute(compute_virial=wl.compute_virial)

        launches_per_step = len(pots)
        n_local = n_total
        n_bar = float(nl.n_neigh[:state.N].double().mean().item())
        exchange_bytes = 0
        sched = None
    else:
        from azplugins_b200 import slices

        sched = slices.SliceScheduler.from_workload(wl, rank, world, dev, dtype=np.float32,
                                                    buffer=synth.BUFFER, transport=args.transport)
        if not args.no_tune:
            tuned = sched.tune(compute_virial=wl.compute_virial)
        else:
            tuned = []
        step = lambda: sched.step(compute_virial=wl.compute_virial)  # noqa: E731
        launches_per_step = sched.launches_per_step
        n_local = sched.n_local
        n_bar = sched.mean_row_length()
        exchange_bytes = sched.exchange_bytes_per_step()

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()

    # ---- device-resident timing: K steps between two events on the launching stream -----------
    def timed_region():
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(K):
            step()
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    def load_probe():
        # same step, untimed; a fixed count so that every rank runs the same number of steps
        # (the multi-GPU step contains device barriers)
        for _ in range(probe_steps):
            step()
        torch.cuda.synchronize()

    # ~1 s of load from the warm-up's own timing (identical on every rank: all-reduced)
    t0 = time.perf_counter()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    est = torch.tensor([(time.perf_counter() - t0) / 3.0], dtype=torch.float64, device=dev)
    if multi:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    probe_steps = int(min(20000, max(10, 1.0 / max(float(est.item()), 1e-6))))
    ms_total, clocks = timed_with_clocks(timed_region, local_rank, load_probe)
    clocks["sampled_over"] = "timed region + %d untimed repeats of the same step" % probe_steps
    if multi:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / K
    value = n_total / (ms_step * 1e-3)

    # ---- end to end through the public API with host buffers --------------------------------
    if not multi:
        host_pos = torch.empty_like(state.pos, device="cpu").pin_memory()
        host_pos.copy_(state.pos)
        host_force = torch.empty_like(pots[0]._force, device="cpu").pin_memory()
        host_virial = torch.empty_like(pots[0]._virial, device="cpu").pin_memory()
        h2d = host_pos.numel() * host_pos.element_size()
        d2h = (host_force.numel() * host_force.element_size()) * len(pots)
        if wl.compute_virial:
            d2h += host_virial.numel() * host_virial.element_size() * len(pots)

        def e2e_step():
            # the call a user makes: positions from pinned host memory in, per-particle forces
            # (+ virials) delivered to pinned host memory (row chunks, copies overlapped)
            state.pos.copy_(host_pos, non_blocking=True)
            for p in pots:
                p.compute_to_host(host_force, host_virial if wl.compute_virial else None)
    else:
        h2d, d2h = sched.e2e_bytes(wl.compute_virial)
        e2e_step = lambda: sched.e2e_step(compute_virial=wl.compute_virial)  # noqa: E731

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    if multi:
        t = torch.tensor([e2e_ms, e2e_wall_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms, e2e_wall_ms = float(t[0].item()), float(t[1].item())
    e2e_value = n_total / (max(e2e_ms, e2e_wall_ms) / K * 1e-3)

    # ---- roofline of the dominant kernel (per-rank: it processes n_local rows per launch) ------
    # one launch per potential; time = device time of the step / launches (steps are kernel-only
    # at N = 1; at N > 1 the step also holds the exchange, so the kernel is timed separately)
    if multi:
        kern_ms = sched.time_kernels(K, compute_virial=wl.compute_virial)
        t = torch.tensor([kern_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kern_ms = float(t.item())
    else:
        kern_ms = ms_step
    # per potential: fixed arrays + the neighbour list once; a workload with two
    # potentials (C3) streams the list twice, and the step time covers both launches
    n_pot = len(wl.potentials)
    bytes_fixed = wl.bytes_per_particle - 4.0 * wl.n_bar if wl.bytes_per_particle else 44.0
    alg_bytes = (bytes_fixed + 4.0 * n_bar) * n_local * n_pot
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": NCU_TRAFFIC_BYTES.get(args.workload),
                "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                "kernel_ms": kern_ms, "algorithmic_bytes_per_step": alg_bytes,
                "bytes_per_particle": (bytes_fixed + 4.0 * n_bar) * n_pot,
                "mean_row_length": n_bar, "launches_per_step": launches_per_step}
    if args.workload in NCU_WARP_INSTRUCTIONS and not multi and clocks.get("sm_mhz"):
        # informational co-limiter: one warp instruction per scheduler per clock
        floor_ms = (NCU_WARP_INSTRUCTIONS[args.workload] * (n_local / 1.0e6)
                    / (SM_COUNT * SCHEDULERS_PER_SM * clocks["sm_mhz"] * 1e6) * 1e3)
        roofline["issue_floor_ms"] = floor_ms
        roofline["issue_frac"] = floor_ms / kern_ms

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "N": n_total, "particles_per_gpu": n_local,
                       "types": len(wl.types), "mean_row_length": n_bar,
                       "compute_virial": bool(wl.compute_virial),
                       "launch_shape_block_tpp_ms": tuned,
                       "l2_policy": "inputs larger than L2: the %.0f MB neighbour list streams "
                                    "from HBM every step; positions stay L2-resident by design"
                                    % (4e-6 * n_bar * n_local),
                       "parallelism": "1 GPU" if not multi else
                       "%d particle slices, %s (%.1f MB/step/rank)"
                       % (world, "halo pushed over NVLink peer memory" if sched.transport == "peer"
                          else "NCCL halo exchange", exchange_bytes / 1e6)},
            "roofline": roofline, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": max(e2e_ms, e2e_wall_ms) / K},
            "gpu_launches": launches_per_step * K}

    if rank == 0 and not multi and not args.no_cpu_baseline:
        _, info, _ = cpu_oracle_rate(wl, target_seconds=8.0)
        line["cpu_baseline"] = info
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        emit(line)
    if multi:
        dist.destroy_process_group()


def main():
    # the contract is ONE JSON line on stdout: libraries (NCCL prints its version banner there)
    # are kept off it by pointing fd 1 at stderr until the line is written
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    if args.impl == "reference":
        # the CPU arm uses every host core it may run on; torchrun exports OMP_NUM_THREADS=1,
        # which must be overridden before the OpenMP runtime is loaded
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        os.environ["OMP_NUM_THREADS"] = str(cores)
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
