#!/usr/bin/env python
"""Neighbour-list build time (row f1) and the fp64 force kernel on the C2 fluid.
    python tools/nlist_bench.py [N]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import azplugins_b200 as az
from azplugins_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
wl = synth.config2(N=N)
for dtype in (np.float32, np.float64):
    state = wl.make_state(dtype=dtype)
    nl = az.nlist.Cell(buffer=synth.BUFFER)
    (pot,) = wl.make_potentials(nl)
    pot.attach(state)
    nl.build(state)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        nl.build(state)
    torch.cuda.synchronize()
    t_build = (time.perf_counter() - t0) / 5
    nl.freeze()
    b, t, ms = pot.tune_kernel_parameters(compute_virial=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        pot.compute(compute_virial=True)
    e0.record()
    for _ in range(20):
        pot.compute(compute_virial=True)
    e1.record()
    torch.cuda.synchronize()
    print("%s N=%d: neighbour-list build %.2f ms (%d entries, wall clock incl. host logic); force kernel "
          "%.4f ms at (block %d, tpp %d)" % (np.dtype(dtype).name, N, 1e3 * t_build, nl.size,
                                             e0.elapsed_time(e1) / 20, b, t), flush=True)
    del state, nl, pot
    torch.cuda.empty_cache()
