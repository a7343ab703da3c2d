#!/usr/bin/env python
"""MD loop (velocity Verlet + pair forces + neighbour-list rebuilds) through the public API:
time steps per second and particle-steps per second, end to end on the device.
    python tools/md_loop.py [C1|C2] [N] [steps] [rebuild_check_delay] [graph]"""
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import azplugins_b200 as az
from azplugins_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C1"
N = int(sys.argv[2]) if len(sys.argv) > 2 else {"C1": 32000, "C2": 1000000}[cfg]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
delay = int(sys.argv[4]) if len(sys.argv) > 4 else 1      # nlist rebuild_check_delay
graph = len(sys.argv) > 5 and sys.argv[5] == "graph"      # replay the un-checked steps from a CUDA graph
rng = np.random.default_rng(3)
rho = {"C1": 0.8, "C2": 0.5}[cfg]
xyz, L = synth.jittered_lattice(N, rho, rng, jitter=0.05)
perm = synth.morton_order(xyz, L)
xyz = xyz[perm]
v = rng.standard_normal((N, 3))
v -= v.mean(axis=0)
nl = az.nlist.Cell(buffer=0.4, rebuild_check_delay=delay)
if cfg == "C1":
    types, typeid = ["A"], np.zeros(N, dtype=np.uint32)
    pot = az.pair.PerturbedLennardJones(nlist=nl, default_r_cut=3.0, mode="shift")
    pot.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0, attraction_scale_factor=0.5)
else:
    types, typeid = ["A", "B"], (rng.random(N) < 0.5).astype(np.uint32)
    pot = az.pair.ExpandedYukawa(nlist=nl, default_r_cut=3.5, mode="shift")
    pot.params[("A", "A")] = dict(epsilon=1.0, kappa=1.0, delta=0.0)
    pot.params[("A", "B")] = dict(epsilon=2.0, kappa=1.2, delta=0.15)
    pot.params[("B", "B")] = dict(epsilon=3.0, kappa=1.5, delta=0.3)
state = az.State(az.Box.cube(L), types, xyz, typeid=typeid, velocity=v, dtype=np.float32)
ig = az.md.Integrator(dt=0.002, forces=[pot]).attach(state)
ig.run(20, graph=graph)
torch.cuda.synchronize()
e0 = ig.kinetic_energy() + ig.potential_energy()
b0 = nl.num_builds
t0 = time.perf_counter()
ig.run(steps, graph=graph)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
e1 = ig.kinetic_energy() + ig.potential_energy()
print("[check delay %d, %s] " % (delay, "graph replay" if graph else "eager"), end="")
print("%s N=%d: %d steps in %.3f s = %.1f time steps/s = %.3e particle-steps/s; %d neighbour-list "
      "builds (%d reused the row capacities); total energy %.6g -> %.6g (fp32)"
      % (cfg, N, steps, dt, steps / dt, N * steps / dt, nl.num_builds - b0,
         getattr(nl, "num_reused", 0), e0, e1))
