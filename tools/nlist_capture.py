#!/usr/bin/env python
"""One rebuild of the neighbour list (bin + sort + fill, capacities reused) of the C2 fluid at
N = 1 M inside a cudaProfilerStart/Stop bracket, for
    ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/<tag> \
        python tools/nlist_capture.py
Also prints the device time of the same rebuild measured with CUDA events before the bracket
(outside the profiler's replays), split into the bin (azp_nlist_bin) and the rows pass."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import azplugins_b200 as az
from azplugins_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
wl = synth.config2(N=N)
state = wl.make_state(dtype=np.float32)
nl = az.nlist.Cell(buffer=synth.BUFFER)
for p in wl.make_potentials(nl):
    p.attach(state)
for _ in range(3):
    nl.build(state)  # count + fill, then capacities reused
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    nl.build(state)
e1.record()
torch.cuda.synchronize()
print(json.dumps(dict(N=N, rebuild_ms=e0.elapsed_time(e1) / 10, entries=int(nl.n_neigh.sum().item()),
                      reused=nl.num_reused)), flush=True)
torch.cuda.cudart().cudaProfilerStart()
nl.build(state)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
