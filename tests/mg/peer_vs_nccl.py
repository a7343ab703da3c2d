"""Launched by tests/test_multi_gpu.py under torchrun (one rank per GPU): the NVLink peer-memory
halo push must give bit-identical ghosts and forces to the NCCL exchange, and both must equal a
single-GPU evaluation of the same global system on the rows of this rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import azplugins_b200 as az  # noqa: E402
from azplugins_b200 import slices, synth  # noqa: E402


def main():
    cfg, n_per = sys.argv[1], int(sys.argv[2])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    wl = synth.CONFIGS[cfg](N=n_per * world)
    virial = True
    out = {}
    for transport in ("nccl", "peer"):
        s = slices.SliceScheduler.from_workload(wl, rank, world, dev, transport=transport)
        for p in s.pots:
            p.kernel_parameters = (128, 1)  # same summation order whatever the row partition
        # scramble the ghosts so that only a working exchange can restore them
        for a in s.exchange_arrays:
            a[s.n_local:].fill_(1.0e3)
        for _ in range(2):
            s.step(compute_virial=virial)
        torch.cuda.synchronize()
        out[transport] = dict(ghost=[a[s.n_local:].clone() for a in s.exchange_arrays],
                              force=[p._force.clone() for p in s.pots],
                              virial=[p._virial.clone() for p in s.pots],
                              torque=[p._torque.clone() for p in s.pots])
        lo, hi = s.plan.lo, s.plan.hi
        del s
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers

    ok = True
    # ghosts: the two transports must deliver the same bytes. Forces: bit-identical too, except
    # for the DPD family, whose deferred-accept queue pops a row's neighbours in an order that
    # depends on the other rows of the warp (the NCCL path splits rows into interior/boundary
    # launches): there only the fp32 summation order may differ.
    exact = cfg != "C4"
    for key in ("ghost", "force", "virial", "torque"):
        for x, y in zip(out["nccl"][key], out["peer"][key]):
            if key == "ghost" or exact:
                same = torch.equal(x, y)
            else:
                xn, yn = x.cpu().numpy(), y.cpu().numpy()
                if key == "virial":
                    same = helpers.virial_rel_err(xn.T, yn) < 1e-6
                else:
                    same = helpers.per_particle_rel_err(xn[:, :3], yn[:, :3]) < 1e-5
            ok = ok and same
            if not same:
                print("rank %d: %s differs between transports (max %g)" % (rank, key, (x - y).abs().max().item()))
    # single-GPU evaluation of the whole system, rows [lo, hi)
    state = wl.make_state(dtype=np.float32, device=dev)
    nl = az.nlist.Cell(buffer=synth.BUFFER)
    pots = wl.make_potentials(nl)
    for p, f_mg, v_mg in zip(pots, out["peer"]["force"], out["peer"]["virial"]):
        p.attach(state).compute(compute_virial=virial)
        torch.cuda.synchronize()
        f1 = p._force[lo:hi].cpu().numpy()
        fm = f_mg.cpu().numpy()
        err = helpers.per_particle_rel_err(fm[:, :3], f1[:, :3])
        eerr = helpers.total_rel_err(fm[:, 3], f1[:, 3])
        # the same pairs in the same row order; only the launch shape (summation order) may
        # differ from the single-GPU default, so fp32 rounding noise is all that is allowed
        good = err < 1e-5 and eerr < 1e-6
        ok = ok and good
        print("rank %d %s: vs single GPU force err %.2e energy err %.2e %s" % (rank, type(p).__name__, err, eerr, "ok" if good else "FAIL"))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("PEER_VS_NCCL_OK" if int(flag.item()) == 1 else "PEER_VS_NCCL_FAIL")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
