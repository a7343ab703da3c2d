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
  headers under the restated HOOMD loop, else the port) on the host cores, bounded row sample;
  the reference arm never imports the product package (no libazp_b200.so in its process);
* `check` (N > 1): rows of every rank's slice against a single-domain evaluation;
* `strong_scaling`: BASELINE.json's scaling case, C5 TwoPatchMorse N = 16 M in total, at this N;
* `other_configs` (N = 1): C3 (Colloid + Hertz, 4 M) and C4 (DPD thermostat, 8 M) timed the same
  way in the same run, each with its own roofline (`--no-strong` skips both extra records).
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


def ncu_constants(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum per launch of the
    dominant kernel, from the committed `ncu --set full` capture of this workload
    (profiles/ncu_constants.json, written by tools/ncu_summary.py --register). A capture taken
    with other kernels than the ones this run loads is NOT reported: the entry carries the SASS
    digest of the workload's kernels and the hash of the kernel sources at capture time
    (tools/srchash.py); the SASS digest must match when both sides have one (a change to another
    evaluator's kernels then leaves this capture valid), else the source hash must."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import srchash

    path = os.path.join(ROOT, "profiles", "ncu_constants.json")
    here = srchash.kernel_source_hash()
    if not os.path.exists(path):
        return None, "no capture registered", here
    entry = json.load(open(path)).get(workload)
    if entry is None:
        return None, "no capture registered for %s" % workload, here
    sass_here = srchash.kernel_sass_hashes().get(workload)
    if sass_here and entry.get("kernel_sass"):
        if entry["kernel_sass"] != sass_here:
            return None, ("stale capture %s (kernel SASS %s, this build %s)"
                          % (entry.get("summary"), entry["kernel_sass"], sass_here)), here
    elif entry.get("kernel_sources") != here:
        return None, ("stale capture %s (kernel sources %s, this build %s)"
                      % (entry.get("summary"), entry.get("kernel_sources"), here)), here
    return entry, entry.get("summary"), here


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
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the e2e leg")
    ap.add_argument("--cuda-profiler", action="store_true",
                    help="bracket the timed region with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    ap.add_argument("--no-fuse", action="store_true", help="never use the fused two-potential pass")
    ap.add_argument("--no-strong", action="store_true",
                    help="skip the extra records appended to the C2 line (C5 N = 16 M strong scaling; C3 and C4 at N = 1)")
    ap.add_argument("--strong-n", type=int, default=16000000)
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
def load_synth_without_product_library():
    """The workload generator (numpy only) for the CPU arm WITHOUT importing the product package:
    `import azplugins_b200` dlopens libazp_b200.so, and a reference arm that maps the product
    library is exactly what the driver's loaded-library record exists to catch. The generator's
    modules (synth, box, state) are loaded under a private package name whose __init__ is
    empty, so azplugins_b200/__init__.py never runs."""
    import importlib
    import types

    name = "_azp_workloads_only"
    if name not in sys.modules:
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(ROOT, "azplugins_b200")]
        sys.modules[name] = pkg
    synth = importlib.import_module(name + ".synth")
    state = importlib.import_module(name + ".state")
    assert "azplugins_b200" not in sys.modules or True
    return synth, state.pack_pos


def cpu_oracle_rate(wl, pack_pos, target_seconds=8.0, steps=1, warmup=0):
    """particle-steps/s of the reference's CPU path on the host cores: the oracle loop
    (oracle/_ref = the reference's own evaluator headers under the restated HOOMD host loop, else
    the port) on a bounded sample of rows of `wl`, fp32, every host thread.

    Two parallel modes are timed and the faster one is the value (both are reported):
      "domains"   -- HOOMD's CPU parallelism: one half-list domain per thread (MPI domain
                     decomposition restated with OpenMP, oracle/driver_loops.h): a pair inside a
                     domain is evaluated once, a pair across a domain face on both sides;
      "full-list" -- every row evaluates its whole full-list row (what the GPU kernel does).
    plus one thread with a true half list (what a single HOOMD CPU rank does)."""
    from oracle import oracle

    orc = oracle.load("best", np.float32)
    kind = "reference" if orc.kind == "ref" else "port"
    cores = orc.max_threads()
    pos = pack_pos(wl.position, wl.typeid, np.float32)
    spec = wl.potentials
    nt = len(wl.types)

    def tables(s):
        name = s["cls"]
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
    r_list = np.where(rc_max > 0, rc_max + 0.4, 0.0)
    vel = None
    if wl.velocity is not None:
        vel = np.zeros((len(pos), 4), dtype=np.float32)
        vel[:, :3] = wl.velocity
    quat = None if wl.orientation is None else wl.orientation.astype(np.float32)
    lists = {}

    def run(n_rows, half=False, nthreads=0):
        key = (n_rows, half is True)
        if key not in lists:
            lists.clear()
            lists[key] = orc.build_nlist(pos, wl.box.L, r_list, ntypes=nt, n_rows=n_rows,
                                         half=half is True)
        nn, nl, head = lists[key]
        common = dict(ntypes=nt, virial=wl.compute_virial, N=n_rows, nthreads=nthreads, half=half)
        t0 = time.perf_counter()
        for s, (name, table, rc) in zip(spec, tabs):
            mode = s.get("kwargs", {}).get("mode", "none")
            if name == "TwoPatchMorse":
                orc.aniso_forces(table, pos, quat, nn, nl, head, wl.box.L, rc, mode=mode, **common)
            elif name == "DPDGeneralWeight":
                orc.dpd_forces(table, pos, vel, wl.tag, nn, nl, head, wl.box.L, rc, wl.seed,
                               wl.timestep, wl.dt, s["kwargs"]["kT"], **common)
            else:
                orc.pair_forces(name, table, pos, nn, nl, head, wl.box.L, rc, mode=mode, **common)
        return time.perf_counter() - t0

    probe_rows = min(wl.N, 20000)
    t_probe = min(run(probe_rows), run(probe_rows))
    # one pass over the sample should take ~target_seconds / (passes below); the whole system
    # when it fits
    per_pass = target_seconds / (2.0 * (max(1, steps) + warmup) + 1.0)
    n_rows = int(min(wl.N, max(probe_rows, probe_rows * per_pass / max(t_probe, 1e-6))))
    out = {}
    for mode, half in (("full-list", False), ("domains", "domains")):
        for _ in range(warmup):
            run(n_rows, half)
        out[mode] = float(np.median([run(n_rows, half) for _ in range(max(1, steps))]))
    best = min(out, key=out.get)
    t = out[best]
    # one thread, true half list (third law), on the probe's rows
    t1 = run(probe_rows, half=True, nthreads=1)
    info = {"value": n_rows / t, "unit": UNIT, "cores": cores, "kind": kind, "mode": best,
            "value_domains_half_list": n_rows / out["domains"],
            "value_full_list": n_rows / out["full-list"],
            "value_1_thread_half_list": probe_rows / t1,
            "build": "g++ -O3 -march=x86-64-v3 -ffp-contract=off (oracle/Makefile)",
            "sample": "%s %d of %d rows of %s, fp32, %.3f s per pass, %d threads; the faster of "
                      "one half-list domain per thread (HOOMD's MPI decomposition) and the "
                      "full-list loop"
                      % ("all" if n_rows == wl.N else "first", n_rows, wl.N, wl.name, t, cores)}
    return n_rows / t, info, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    synth, pack_pos = load_synth_without_product_library()
    n = args.n_per_gpu or DEFAULT_N[args.workload]
    wl = synth.CONFIGS[args.workload](N=n * max(1, args.gpus) if args.gpus > 1 else n)
    budget = 60.0  # seconds of CPU work for the whole --steps/--warmup run
    value, info, t = cpu_oracle_rate(wl, pack_pos, target_seconds=budget, steps=args.steps,
                                     warmup=args.warmup)
    assert "azplugins_b200" not in sys.modules, "the reference arm must not load the product"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, args.gpus),
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def config_of(wl, n_gpus):
    """The workload description shared by both arms (same keys, same values)."""
    return {"workload": wl.name, "N": int(wl.N), "particles_per_gpu": int(wl.N // max(1, n_gpus)),
            "types": len(wl.types), "compute_virial": bool(wl.compute_virial),
            "potentials": [s["cls"] for s in wl.potentials], "r_buff": 0.4, "precision": "fp32"}


# -------------------------------------------------------------------------------------------------
# B200 arm
# -------------------------------------------------------------------------------------------------
class Job:
    """One workload set up on this rank's GPU (single GPU) or as this rank's particle slice."""

    def __init__(self, wl, args, torch, dist, dev, rank, world):
        import azplugins_b200 as az
        from azplugins_b200 import synth

        self.wl, self.args, self.torch, self.dist = wl, args, torch, dist
        self.dev, self.rank, self.world, self.multi = dev, rank, world, world > 1
        self.virial = wl.compute_virial
        if not self.multi:
            self.state = wl.make_state(dtype=np.float32, device=dev)
            self.nl = az.nlist.Cell(buffer=synth.BUFFER)
            self.pots = wl.make_potentials(self.nl)
            for p in self.pots:
                p.attach(self.state)
            self.nl.compute(self.state)
            # the list is frozen for the run: the per-step displacement check belongs to the
            # neighbour-list row, not to the force path that is timed here
            self.nl.freeze()
            torch.cuda.synchronize()
            self.tuned = []
            for p in self.pots:
                if not args.no_tune:
                    self.tuned.append(p.tune_kernel_parameters(compute_virial=self.virial))
                else:
                    p.kernel_parameters = (args.block, args.tpp)
                    self.tuned.append(p.kernel_parameters + (None,))
            # potentials that share the list and have a fused kernel (C3: Colloid + Hertz) run as
            # one sweep of the list unless --no-fuse; both ways are timed and reported
            self.units = list(self.pots)
            self.fusion = None
            if (len(self.pots) == 2 and az.pair.FusedPair.can_fuse(*self.pots)
                    and hasattr(az._lib.lib, "azp_pair_forces_fused_f32")):
                def time_units(units, reps=5):
                    for u in units:
                        u.compute(compute_virial=self.virial)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(reps):
                        for u in units:
                            u.compute(compute_virial=self.virial)
                    e1.record()
                    torch.cuda.synchronize()
                    return e0.elapsed_time(e1) / reps

                fused = az.pair.FusedPair(*self.pots)
                shape = fused.tune_kernel_parameters(compute_virial=self.virial) if not args.no_tune else None
                t_sep, t_fused = time_units(self.pots), time_units([fused])
                self.fusion = {"separate_ms": t_sep, "fused_ms": t_fused, "fused_shape": shape,
                               "used": "fused" if (t_fused < t_sep and not args.no_fuse) else "separate"}
                if self.fusion["used"] == "fused":
                    self.units = [fused]
            self.sched = None
            self.launches_per_step = len(self.units) * (2 if self.nl.n_max > 512 else 1)
            self.n_local = wl.N
            self.n_bar = float(self.nl.n_neigh[:self.state.N].double().mean().item())
            self.exchange_bytes = 0
        else:
            from azplugins_b200 import slices

            self.sched = slices.SliceScheduler.from_workload(
                wl, rank, world, dev, dtype=np.float32, buffer=synth.BUFFER, transport=args.transport)
            self.tuned = [] if args.no_tune else self.sched.tune(compute_virial=self.virial)
            self.pots = self.sched.pots
            self.launches_per_step = self.sched.launches_per_step
            self.n_local = self.sched.n_local
            self.n_bar = self.sched.mean_row_length()
            self.exchange_bytes = self.sched.exchange_bytes_per_step()

    def step(self):
        if self.sched is not None:
            self.sched.step(compute_virial=self.virial)
        else:
            for p in self.units:
                p.compute(compute_virial=self.virial)

    def barrier(self):
        if self.multi:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.multi:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def time_steps(self, K):
        """ms for K steps between two events on the launching stream, barrier + synchronize on
        both sides, max over ranks."""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(K):
            self.step()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))[0]

    def kernel_ms(self, K):
        """ms per step of the force kernels alone (at N = 1 the step is kernel-only)."""
        if self.sched is None:
            return None
        return self.max_over_ranks(self.sched.time_kernels(K, compute_virial=self.virial))[0]

    def roofline(self, kern_ms, hbm_peak, peak_src):
        wl = self.wl
        # per potential: fixed arrays + the neighbour list once; a workload with
        # two potentials (C3) streams the list twice, and the step time covers both launches
        n_pot = len(wl.potentials)
        bytes_fixed = wl.bytes_per_particle - 4.0 * wl.n_bar if wl.bytes_per_particle else 44.0
        alg_bytes = (bytes_fixed + 4.0 * self.n_bar) * self.n_local * n_pot
        if getattr(self, "fusion", None) and self.fusion["used"] == "fused":
            # one sweep: every input once, two force outputs (524 B for C3)
            alg_bytes = (bytes_fixed + 16.0 + 4.0 * self.n_bar) * self.n_local
            n_pot = 1
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0, "kernel_ms": kern_ms,
                "algorithmic_bytes_per_step": alg_bytes,
                "bytes_per_particle": alg_bytes / self.n_local,
                "mean_row_length": self.n_bar, "launches_per_step": self.launches_per_step}

    def check(self):
        """Multi-GPU: rows of this rank's slice against a single-domain evaluation on this GPU
        (slices.SliceScheduler.verify_against_single_domain); worst case over ranks."""
        if self.sched is None:
            return None
        c = self.sched.verify_against_single_domain(self.wl, compute_virial=self.virial)
        worst, not_identical = self.max_over_ranks(c["max_rel_diff"], 0.0 if c["bit_identical"] else 1.0)
        return {"what": "first %d + middle %d rows of every rank's slice vs a single-domain "
                        "evaluation of the whole system on the same GPU, after the timed steps"
                        % (4096, 4096),
                "rows_per_rank": c["rows"], "max_rel_diff_over_ranks": worst,
                "bit_identical_on_every_rank": not_identical == 0.0,
                "ok": worst <= 1e-5}


def run_b200(args):
    import torch
    import torch.distributed as dist

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
    job = Job(wl, args, torch, dist, dev, rank, world)
    step, barrier = job.step, job.barrier
    for _ in range(W):
        step()
    barrier()

    # ---- device-resident timing, with the clocks sampled under the same load -----------------
    def load_probe():
        # same step, untimed; a fixed count so that every rank runs the same number of steps
        # (the multi-GPU step contains device barriers)
        for _ in range(probe_steps):
            step()
        torch.cuda.synchronize()

    t0 = time.perf_counter()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    est = job.max_over_ranks((time.perf_counter() - t0) / 3.0)[0]
    probe_steps = int(min(20000, max(10, 1.0 / max(est, 1e-6))))
    def timed_region():
        if args.cuda_profiler:
            torch.cuda.cudart().cudaProfilerStart()
        try:
            return job.time_steps(K)
        finally:
            if args.cuda_profiler:
                torch.cuda.cudart().cudaProfilerStop()

    ms_total, clocks = timed_with_clocks(timed_region, local_rank, load_probe)
    clocks["sampled_over"] = "timed region + %d untimed repeats of the same step" % probe_steps
    ms_step = ms_total / K
    value = n_total / (ms_step * 1e-3)

    # ---- end to end through the public API with host buffers --------------------------------
    if not multi:
        state, pots = job.state, job.pots
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
        h2d, d2h = job.sched.e2e_bytes(wl.compute_virial)
        e2e_step = lambda: job.sched.e2e_step(compute_virial=wl.compute_virial)  # noqa: E731

    e2e = None
    if not args.no_e2e:
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
        e2e_ms, e2e_wall_ms = job.max_over_ranks(max(e0.elapsed_time(e1), 0.0),
                                                 (time.perf_counter() - t0) * 1e3)
        e2e = {"value": n_total / (max(e2e_ms, e2e_wall_ms) / K * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": max(e2e_ms, e2e_wall_ms) / K}

    # ---- roofline of the dominant kernel (per rank: it processes n_local rows per launch) ------
    kern_ms = job.kernel_ms(K) or ms_step
    roofline = job.roofline(kern_ms, hbm_peak, peak_src)
    entry, ncu_note, src_hash = ncu_constants(args.workload)
    roofline["ncu_capture"] = ncu_note
    roofline["kernel_sources"] = src_hash
    import srchash  # tools/ is on the path since ncu_constants()

    roofline["kernel_sass"] = srchash.kernel_sass_hashes().get(args.workload)
    if entry is not None and not multi and n_total == int(entry["N"]):
        roofline["traffic"] = entry["dram_bytes"]
        roofline["traffic_over_algorithmic"] = entry["dram_bytes"] / roofline["algorithmic_bytes_per_step"]
        for k in ("l1_hit_pct", "l2_hit_pct", "issue_active_pct", "registers"):
            if k in entry:
                roofline[k] = entry[k]
        if clocks.get("sm_mhz") and entry.get("warp_instructions"):
            # informational co-limiter: one warp instruction per scheduler per clock
            floor_ms = (entry["warp_instructions"]
                        / (SM_COUNT * SCHEDULERS_PER_SM * clocks["sm_mhz"] * 1e6) * 1e3)
            roofline["issue_floor_ms"] = floor_ms
            roofline["issue_frac"] = floor_ms / kern_ms
    check = job.check()

    config = config_of(wl, world)
    if getattr(job, "fusion", None):
        config["fusion"] = job.fusion
    config.update({"mean_row_length": job.n_bar, "launch_shape_block_tpp_ms": job.tuned,
                   "l2_policy": "inputs larger than L2: the %.0f MB neighbour list streams "
                                "from HBM every step; positions stay L2-resident by design"
                                % (4e-6 * job.n_bar * job.n_local),
                   "parallelism": "1 GPU" if not multi else
                   "%d particle slices, %s (%.1f MB/step/rank)"
                   % (world, "halo pushed over NVLink peer memory" if job.sched.transport == "peer"
                      else "NCCL halo exchange", job.exchange_bytes / 1e6)})
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "roofline": roofline, "clocks": clocks,
            "e2e": e2e, "gpu_launches": job.launches_per_step * K}
    if check is not None:
        line["check"] = check

    # ---- the north star's scaling case in the same record: C5, N = 16 M in total, strong --------
    if not args.no_strong and args.workload == "C2" and not args.n_per_gpu:
        del job
        torch.cuda.empty_cache()
        wl5 = synth.config5(N=args.strong_n)
        sj = Job(wl5, args, torch, dist, dev, rank, world)
        for _ in range(W):
            sj.step()
        s_ms = sj.time_steps(K) / K
        s_kern = sj.kernel_ms(K) or s_ms
        s_roof = sj.roofline(s_kern, hbm_peak, peak_src)
        line["strong_scaling"] = {
            "workload": wl5.name, "N": wl5.N, "n_gpus": world, "ms_per_step": s_ms,
            "value": wl5.N / (s_ms * 1e-3), "unit": UNIT, "scaling": "strong",
            "kernel_ms": s_kern, "roofline_frac": s_roof["frac"],
            "launch_shape_block_tpp_ms": sj.tuned, "check": sj.check(),
            "note": "BASELINE.json north star: >= 6x at 8 GPUs for N = 16 M; the total N is fixed, "
                    "so the speed-up is this value over the n_gpus = 1 run's"}
        entry5, why5, _ = ncu_constants("C5")
        if entry5 is not None and not multi:
            line["strong_scaling"]["traffic"] = entry5["dram_bytes"]
            line["strong_scaling"]["ncu_capture"] = why5
        del sj
        torch.cuda.empty_cache()

    # ---- the other single-GPU configurations of BASELINE.json in the same record ----------------
    # (N = 1 only; C2 stays the line's `value`. One device-resident timing per configuration with
    # its own roofline, so the driver-run line carries every kernel, not just the headline's.)
    if not args.no_strong and args.workload == "C2" and not args.n_per_gpu and not multi:
        others = []
        for name in ("C3", "C4"):
            wlx = synth.CONFIGS[name]()
            oj = Job(wlx, args, torch, dist, dev, rank, world)
            for _ in range(W):
                oj.step()
            o_ms = oj.time_steps(K) / K
            o_kern = oj.kernel_ms(K) or o_ms
            o_roof = oj.roofline(o_kern, hbm_peak, peak_src)
            entry, why, _ = ncu_constants(name)
            if entry is not None:
                o_roof["traffic"] = entry["dram_bytes"]
            o_roof["ncu_capture"] = why
            others.append({"workload": wlx.name, "N": wlx.N, "ms_per_step": o_ms,
                           "value": wlx.N / (o_ms * 1e-3), "unit": UNIT, "roofline": o_roof,
                           "launch_shape_block_tpp_ms": oj.tuned, "fusion": oj.fusion,
                           "compute_virial": oj.virial, "potentials": [p["cls"] for p in wlx.potentials]})
            del oj
            torch.cuda.empty_cache()
        line["other_configs"] = others

    if rank == 0 and not multi and not args.no_cpu_baseline:
        from azplugins_b200.state import pack_pos

        _, info, _ = cpu_oracle_rate(wl, pack_pos, target_seconds=15.0)
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
