#!/usr/bin/env python
"""Which stream bounds a force kernel? Times the product kernel (pinned launch shape) on the real
workload and on three doctored neighbour lists that each remove ONE memory stream while keeping
the instruction stream and the row lengths:

  real      the builder's list
  hot_list  every row reads the list storage of row (r mod 4096): the 4-byte index stream comes
            from L1/L2 instead of HBM (indices of 32 consecutive rows still point at 32
            consecutive neighbourhoods, so the gather locality is the real one)
  xor1      real list storage, every entry = r xor 1: the position gathers are perfectly local
  both      hot_list + xor1: what is left is instruction issue and the row prologues

    python tools/diag_streams.py --workloads C2 C4:4000000 C5:4000000
The results of the doctored runs are meaningless as forces; only their time is used."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def time_ms(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    import azplugins_b200 as az
    from azplugins_b200 import synth

    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", nargs="+", default=["C2"])
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rows = []
    for spec in a.workloads:
        name, _, n = spec.partition(":")
        wl = synth.CONFIGS[name](N=int(n)) if n else synth.CONFIGS[name]()
        state = wl.make_state(dtype=np.float32, device="cuda:0")
        nl = az.nlist.Cell(buffer=synth.BUFFER)
        pot = wl.make_potentials(nl)[0]
        pot.attach(state)
        pot.compute(compute_virial=wl.compute_virial)
        nl.freeze()
        shape = pot.tune_kernel_parameters(compute_virial=wl.compute_virial)
        fn = lambda: pot.compute(compute_virial=wl.compute_virial)  # noqa: E731
        N = state.N
        real_head = nl.head_list.clone()
        real_list = nl.nlist.clone()
        r = torch.arange(N, device=nl.head_list.device, dtype=torch.int64)
        hot_head = real_head[r % 4096]
        # n_neigh of the borrowed row may exceed this row's: clamp n_neigh to the smaller one
        real_nn = nl.n_neigh.clone()
        hot_nn = torch.minimum(real_nn, real_nn[r % 4096])
        # entry = row xor 1 for every slot of every row (row of a slot by searchsorted on heads)
        slot = torch.arange(real_list.numel(), device=real_list.device, dtype=torch.int64)
        row_of = torch.clamp(torch.searchsorted(real_head, slot, right=True) - 1, 0, N - 1)
        xor_list = (row_of ^ 1).clamp_(max=N - 1).to(real_list.dtype)
        del slot, row_of
        out = dict(workload=spec, shape=shape)
        for label, head, nn, lst in (("real", real_head, real_nn, real_list),
                                     ("hot_list", hot_head, hot_nn, real_list),
                                     ("xor1", real_head, real_nn, xor_list),
                                     ("both", hot_head, hot_nn, xor_list)):
            nl.head_list.copy_(head)
            nl.n_neigh.copy_(nn)
            nl.nlist.copy_(lst)
            out[label + "_ms"] = time_ms(fn)
            out[label + "_entries"] = int(nn.sum().item())
        rows.append(out)
        print(json.dumps(out), flush=True)
        del state, nl, pot
        torch.cuda.empty_cache()
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
