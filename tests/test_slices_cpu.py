"""CPU, world_size 2 over gloo: the multi-GPU particle-slice scheduler's host logic
(azplugins_b200/slices.py) -- partition, ghost discovery, index remapping, the negotiated
point-to-point halo exchange, interior/boundary row classes. The forces of each rank's local+ghost
arrays (computed here by the CPU oracle, which tests may use) must equal the rows of the
single-domain result."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle


def test_partition_bounds():
    from azplugins_b200 import slices

    b = slices.partition_bounds(1000003, 8)
    assert b[0] == 0 and b[-1] == 1000003 and (np.diff(b) > 0).all()
    assert np.diff(b).max() - np.diff(b).min() <= 1
    w = np.ones(1000)
    w[:10] = 100.0  # ten long rows at the front
    bw = slices.partition_bounds_weighted(w, 4)
    sums = [w[bw[r]:bw[r + 1]].sum() for r in range(4)]
    assert bw[0] == 0 and bw[-1] == 1000 and max(sums) / min(sums) < 1.3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out, weighted=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from azplugins_b200 import slices, synth

        dtype = np.float64
        wl = synth.config5(N=4096)  # TwoPatchMorse: positions AND orientations are exchanged
        orc = oracle.load("port", dtype)
        pos = oracle.make_pos(wl.position, wl.typeid, dtype)
        quat = wl.orientation.astype(dtype)
        L = wl.box.L
        r_cut, buf = 1.6, 0.4
        nn, nl, head = orc.build_nlist(pos, L, r_cut + buf)
        if weighted:
            # balance the sum of (skewed) row weights instead of the row count, as
            # SliceScheduler.from_workload(balance="neighbors") does with n_neigh
            w = nn.astype(np.float64) * np.where(np.arange(wl.N) < wl.N // 4, 4.0, 1.0)
            bounds = slices.partition_bounds_weighted(w, world)
            assert abs(int(bounds[1]) - wl.N // 2) > wl.N // 10
        else:
            bounds = slices.partition_bounds(wl.N, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        # rows of this slice with global indices (what the GPU builder returns for rows=(lo,hi))
        h0 = int(head[lo])
        h1 = int(head[hi]) if hi < wl.N else len(nl)
        plan = slices.SlicePlan(rank, world, bounds,
                                torch.from_numpy(nn[lo:hi].view(np.int32).copy()),
                                torch.from_numpy((head[lo:hi] - h0).view(np.int64).copy()),
                                torch.from_numpy(nl[h0:h1].view(np.int32).copy()))
        plan.negotiate()
        assert plan.n_local == hi - lo and plan.n_ghost > 0
        assert len(plan.interior_rows) + len(plan.boundary_rows) == plan.n_local
        assert len(plan.interior_rows) > 0 and len(plan.boundary_rows) > 0
        # ghosts are exactly the non-owned particles referenced by the rows
        ids = plan.ghost_ids.numpy()
        assert ((ids < lo) | (ids >= hi)).all() and (np.diff(ids) > 0).all()
        # interior rows reference no ghost
        nloc = plan.nlist_local.numpy()
        hl = (head[lo:hi] - h0).astype(np.int64)
        for r in plan.interior_rows.numpy()[:200]:
            assert (nloc[hl[r]:hl[r] + nn[lo + r]] < plan.n_local).all()
        for r in plan.boundary_rows.numpy()[:200]:
            assert (nloc[hl[r]:hl[r] + nn[lo + r]] >= plan.n_local).any()
        # local + ghost arrays; ghosts start as garbage and must be filled by the exchange
        gl = np.concatenate([np.arange(lo, hi), ids])
        lpos = torch.from_numpy(pos[gl].copy())
        lquat = torch.from_numpy(quat[gl].copy())
        lpos[plan.n_local:] = 777.0
        lquat[plan.n_local:] = 777.0
        halo = slices.HaloExchange(plan)
        halo([lpos, lquat])
        assert np.array_equal(lpos.numpy(), pos[gl]) and np.array_equal(lquat.numpy(), quat[gl])
        assert halo.bytes_per_step([lpos, lquat]) == plan.n_ghost * 2 * 4 * 8
        # what the NVLink peer push (PeerHalo) addresses: on every receiver, the first ghost row
        # of every owner's segment, and my particles arriving there in the receiver's ghost order
        starts = plan.peer_ghost_start
        assert starts.shape == (world, world)
        assert starts[rank][rank] >= plan.n_local and (starts[rank] >= plan.n_local).all()
        for owner in range(world):
            o, c = int(plan.recv_offsets[owner]), int(plan.recv_counts[owner])
            assert starts[rank][owner] == plan.n_local + o
            if c:
                assert (ids[o:o + c] >= bounds[owner]).all() and (ids[o:o + c] < bounds[owner + 1]).all()
        for peer in range(world):
            if peer != rank:
                # rows I push to `peer` land at starts[peer][rank] ... + send_counts[peer]
                n_peer_local = int(bounds[peer + 1] - bounds[peer])
                assert starts[peer][rank] >= n_peer_local
                assert int(plan.send_counts[peer]) == int(plan.send_idx[peer].numel())
        # a second exchange after the owners moved their particles
        pos2 = pos.copy()
        pos2[:, :3] += 0.01 * np.sin(np.arange(wl.N))[:, None]
        lpos[:plan.n_local] = torch.from_numpy(pos2[lo:hi])
        halo([lpos, lquat])
        assert np.array_equal(lpos.numpy(), pos2[gl])
        # forces of the local rows from local+ghost arrays == rows of the single-domain result
        table = orc.pack_table("TwoPatchMorse", 1, {(0, 0): wl.potentials[0]["params"][("A", "A")]})
        f_loc, t_loc, v_loc = orc.aniso_forces(table, lpos.numpy(), lquat.numpy(), nn[lo:hi], nloc,
                                               hl.astype(np.uint64), L, r_cut, N=plan.n_local)
        f_glob, t_glob, v_glob = orc.aniso_forces(table, pos2, quat, nn, nl, head, L, r_cut)
        assert np.array_equal(f_loc, f_glob[lo:hi]) and np.array_equal(t_loc, t_glob[lo:hi])
        assert np.array_equal(v_loc, v_glob[:, lo:hi])
        out.put((rank, plan.n_local, plan.n_ghost, int(len(plan.interior_rows))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("weighted", [False, True], ids=["by-count", "by-weight"])
def test_two_rank_halo_exchange_and_row_parity(weighted):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, weighted)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    got = sorted(out.get(timeout=5) for _ in range(2))
    assert got[0][1] + got[1][1] == 4096
