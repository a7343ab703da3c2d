#!/usr/bin/env python
"""Neighbour-list build time (row f1) on the C2 fluid (and C4 with --c4): device time of a rebuild
that reuses the row capacities (bin + fill) per lanes-per-row setting, the displacement check,
and the device Morton sort.
    python tools/nlist_bench.py [N] [--c4]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import azplugins_b200 as az
from azplugins_b200 import synth

args = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(args[0]) if args else 1000000
cfg = synth.config4 if "--c4" in sys.argv else synth.config2
wl = cfg(N=N)
out = dict(workload=wl.name, N=N, builds=[])
for tpr in (0, 2, 4, 8, 16, 32):
    state = wl.make_state(dtype=np.float32)
    nl = az.nlist.Cell(buffer=synth.BUFFER, threads_per_row=tpr)
    pots = wl.make_potentials(nl)
    for p in pots:
        p.attach(state)
    nl.build(state)  # count + fill
    nl.build(state)  # capacities reused from here on
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(10):
        nl.build(state)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 10
    row = dict(threads_per_row=tpr, build_ms_device=e0.elapsed_time(e1) / 10, build_ms_wall=1e3 * wall,
               entries=int(nl.n_neigh.sum().item()), slots=int(nl.size), reused=nl.num_reused)
    # the displacement check alone
    e0.record()
    for _ in range(20):
        nl._needs_rebuild(state)
    e1.record()
    torch.cuda.synchronize()
    row["check_ms"] = e0.elapsed_time(e1) / 20
    out["builds"].append(row)
    print(json.dumps(row), flush=True)
    del state, nl, pots
    torch.cuda.empty_cache()
state = wl.make_state(dtype=np.float32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
state.sfc_sort()
e0.record()
for _ in range(5):
    state.sfc_sort()
e1.record()
torch.cuda.synchronize()
out["sfc_sort_ms"] = e0.elapsed_time(e1) / 5
print(json.dumps(dict(sfc_sort_ms=out["sfc_sort_ms"])), flush=True)
if len(args) > 1:
    json.dump(out, open(args[1], "w"), indent=1)
