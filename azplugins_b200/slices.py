"""Multi-GPU particle-slice scheduler: one process per GPU, each owning a
contiguous range of the spatially sorted particles and the neighbour-list rows of that range.

With a full neighbour list every row is independent and writes only its own outputs, so the path
shards by rows with no reduction. What a rank needs from its peers are the positions (and
velocities / orientations) of the particles its rows reference but it does not own -- its
*ghosts*, exactly HOOMD's domain-decomposition picture (rows [0, N) local, indices >= N ghosts,
SURVEY.md Appendix A.2). Per step:

    1. pack the locally owned particles each peer asked for            (comm stream)
    2. exchange them point-to-point over NCCL / NVLink                 (comm stream)
    3. meanwhile evaluate the INTERIOR rows -- rows without ghosts --  (compute stream)
    4. when the halo has landed, evaluate the BOUNDARY rows            (compute stream)

The plan (who owns what, ghost lists, index remapping, interior/boundary split) is pure index
logic on torch tensors and runs on any device/backend; it is built once per neighbour-list
build. Only the force evaluation needs CUDA.
"""

import numpy as np
import torch
import torch.distributed as dist


def partition_bounds(n_total, world):
    """Row bounds of ``world`` contiguous slices with near-equal particle counts (int64[world+1]).
    (Rows of the BASELINE fluids have near-equal length; for skewed rows pass weights.)"""
    return np.array([(n_total * r) // world for r in range(world + 1)], dtype=np.int64)


def partition_bounds_weighted(weights, world):
    """Bounds balancing the sum of ``weights`` (e.g. n_neigh) instead of the particle count."""
    w = np.asarray(weights, dtype=np.float64)
    c = np.concatenate([[0.0], np.cumsum(w)])
    targets = c[-1] * np.arange(world + 1) / world
    b = np.searchsorted(c, targets, side="left")
    b[0], b[-1] = 0, len(w)
    return np.maximum.accumulate(b).astype(np.int64)


class SlicePlan:
    """Index plan of one rank: ghosts, remapped neighbour list, exchange lists, row classes."""

    def __init__(self, rank, world, bounds, n_neigh, head_list, nlist_global):
        """``n_neigh`` / ``head_list`` / ``nlist_global``: rows of particles [lo, hi) with GLOBAL
        neighbour indices (torch tensors, any device)."""
        self.rank, self.world = rank, world
        self.bounds = np.asarray(bounds, dtype=np.int64)
        lo, hi = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.lo, self.hi, self.n_local = lo, hi, hi - lo
        dev = nlist_global.device
        nn = n_neigh.to(torch.int64)
        head = head_list.to(torch.int64)
        size = int(nlist_global.numel())
        # row of every stored entry and whether the entry is valid (k < n_neigh[row])
        cap = torch.diff(head, append=torch.tensor([size], dtype=torch.int64, device=dev))
        row_of = torch.repeat_interleave(torch.arange(self.n_local, device=dev), cap,
                                         output_size=size)
        k = torch.arange(size, device=dev) - head[row_of]
        valid = k < nn[row_of]
        j = nlist_global.to(torch.int64) & 0xFFFFFFFF
        own = (j >= lo) & (j < hi)
        ghost_entry = valid & ~own
        self.ghost_ids = torch.unique(j[ghost_entry])  # sorted global ids
        self.n_ghost = int(self.ghost_ids.numel())
        # local index: own -> j - lo; ghost -> n_local + rank in ghost_ids; padding -> 0
        gpos = torch.searchsorted(self.ghost_ids, j.clamp(min=0)) if self.n_ghost else torch.zeros_like(j)
        j_local = torch.where(own, j - lo, self.n_local + gpos)
        j_local = torch.where(valid, j_local, torch.zeros_like(j_local))
        self.nlist_local = j_local.to(torch.int32)
        self.n_neigh = n_neigh
        self.head_list = head_list
        # interior rows have no ghost entry
        has_ghost = torch.zeros(self.n_local, dtype=torch.bool, device=dev)
        has_ghost[row_of[ghost_entry]] = True
        rows = torch.arange(self.n_local, dtype=torch.int32, device=dev)
        self.interior_rows = rows[~has_ghost].contiguous()
        self.boundary_rows = rows[has_ghost].contiguous()
        # owner of each ghost (owners are contiguous ranges -> segments of the sorted id list)
        b = torch.from_numpy(self.bounds).to(dev)
        owner = torch.searchsorted(b, self.ghost_ids, right=True) - 1
        self.recv_counts = torch.bincount(owner, minlength=world).cpu().numpy().astype(np.int64)
        self.recv_offsets = np.concatenate([[0], np.cumsum(self.recv_counts)])[:-1]
        self.send_idx = None  # filled by negotiate()

    def negotiate(self, group=None):
        """Tell every owner which of its particles this rank needs; learn what to send."""
        wanted = []
        ids = self.ghost_ids.cpu().numpy()
        for r in range(self.world):
            o, c = int(self.recv_offsets[r]), int(self.recv_counts[r])
            wanted.append(ids[o:o + c])
        if self.world == 1:
            self.send_idx = [torch.zeros(0, dtype=torch.int64, device=self.ghost_ids.device)]
            return self
        gathered = [None] * self.world
        dist.all_gather_object(gathered, wanted, group=group)
        # first ghost row (local index) of every owner's segment on every rank: the destination
        # of a peer-memory push (PeerHalo)
        starts = [None] * self.world
        dist.all_gather_object(starts, [int(self.n_local + o) for o in self.recv_offsets], group=group)
        self.peer_ghost_start = np.asarray(starts, dtype=np.int64)  # [receiver][owner]
        dev = self.ghost_ids.device
        self.send_idx = []
        for r in range(self.world):
            need = np.asarray(gathered[r][self.rank], dtype=np.int64)  # global ids rank r wants from me
            assert need.size == 0 or (need.min() >= self.lo and need.max() < self.hi)
            self.send_idx.append(torch.from_numpy(need - self.lo).to(dev))
        self.send_counts = np.array([int(s.numel()) for s in self.send_idx], dtype=np.int64)
        return self


class HaloExchange:
    """Per-step exchange of one or more Scalar4 arrays laid out [local | ghosts]."""

    def __init__(self, plan, group=None):
        self.plan = plan
        self.group = group
        self.peers = [r for r in range(plan.world) if r != plan.rank
                      and (plan.recv_counts[r] > 0 or plan.send_counts[r] > 0)] if plan.world > 1 else []
        self._send_cat = torch.cat([plan.send_idx[r] for r in self.peers]) if self.peers else None
        self._send_off = np.concatenate([[0], np.cumsum([plan.send_counts[r] for r in self.peers])]) \
            if self.peers else None
        self._buf = {}

    def bytes_per_step(self, arrays):
        n = sum(int(self.plan.recv_counts[r]) for r in self.peers)
        return sum(n * a.shape[1] * a.element_size() for a in arrays)

    def _plan_ops(self, arrays):
        """Point-to-point descriptors for `arrays` (buffers are persistent, so the list is built
        once and re-posted every step)."""
        p = self.plan
        ops, packs = [], []
        nsend = int(self._send_off[-1])
        for a in arrays:
            sbuf = torch.empty((nsend, a.shape[1]), dtype=a.dtype, device=a.device)
            packs.append((a, sbuf))
            for pi_, r in enumerate(self.peers):
                s0, s1 = int(self._send_off[pi_]), int(self._send_off[pi_ + 1])
                if s1 > s0:
                    ops.append(dist.P2POp(dist.isend, sbuf[s0:s1], r, group=self.group))
                c = int(p.recv_counts[r])
                if c:
                    o = p.n_local + int(p.recv_offsets[r])
                    ops.append(dist.P2POp(dist.irecv, a[o:o + c], r, group=self.group))
        return ops, packs

    def _ops_for(self, arrays):
        key = tuple(a.data_ptr() for a in arrays)
        if key not in self._buf:
            self._buf.clear()
            self._buf[key] = self._plan_ops(arrays)
        return self._buf[key]

    def pack(self, arrays):
        """Gather the particles the peers asked for into the persistent send buffers (a few
        microseconds: one coalesced kernel of the library on CUDA, index_select on the CPU)."""
        if not self.peers:
            return
        _, packs = self._ops_for(arrays)
        n_local = self.plan.n_local
        for a, sbuf in packs:
            if a.is_cuda:
                from . import kernels

                kernels.gather_rows(a, self._send_cat, sbuf)
            else:
                torch.index_select(a[:n_local], 0, self._send_cat, out=sbuf)

    def exchange(self, arrays):
        """Post the sends/receives of the packed buffers straight into the ghost regions."""
        if not self.peers:
            return
        ops, _ = self._ops_for(arrays)
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def __call__(self, arrays):
        """Update the ghost region of every array in ``arrays`` from the owners (in place)."""
        self.pack(arrays)
        self.exchange(arrays)


class PeerHalo:
    """Halo exchange over NVLink peer memory instead of NCCL send/recv.

    The exchanged arrays live in symmetric memory (``torch.distributed._symmetric_memory``: the
    same allocation on every rank, peer-mapped). One kernel of the library (``azp_push_rows``)
    stores the particles a peer needs straight into that peer's ghost region -- consecutive
    16-byte stores per peer, fire-and-forget over NVLink -- followed by ONE device-side barrier
    on the symmetric-memory signal pads:

        push      my particles -> the peers' ghost regions of buffer set p = step % 2
        barrier   every push has landed; the force kernel may read the ghosts of set p

    The arrays are double-buffered: step t pushes into and computes from set t % 2. A peer last
    read its set t % 2 in the force kernel of step t - 2, which precedes its barrier of step
    t - 1 in stream order -- and this rank passed that barrier before it pushes -- so no second
    barrier is needed to protect the ghosts that are overwritten (round 1 used two barriers per
    step on a single set: 6.6 us each plus the skew they expose). No pack buffer, no second
    stream, no interior/boundary split: the rows run in one launch in natural order right after
    the barrier.
    """

    def __init__(self, plan, arrays, group=None, double_buffer=True):
        import torch.distributed._symmetric_memory as symm_mem

        self.plan = plan
        self.group = group if group is not None else dist.group.WORLD
        world, me = plan.world, plan.rank
        self.peers = [r for r in range(world) if r != me and plan.send_counts[r] > 0]
        dev = arrays[0].device
        n_rows = torch.tensor([arrays[0].shape[0]], dtype=torch.int64, device=dev)
        dist.all_reduce(n_rows, op=dist.ReduceOp.MAX, group=self.group)
        max_rows = int(n_rows.item())
        self._send_cat = torch.cat([plan.send_idx[r] for r in self.peers]) if self.peers else None
        self.sets = []  # per buffer set: (arrays, handles, dst_addr)
        for _ in range(2 if double_buffer else 1):
            sym_arrays, handles, dst_addr = [], [], []
            for a in arrays:
                sym = symm_mem.empty((max_rows, a.shape[1]), dtype=a.dtype, device=dev)
                sym[:a.shape[0]].copy_(a)
                hdl = symm_mem.rendezvous(sym, self.group)
                row_bytes = a.shape[1] * a.element_size()
                addr = []
                for r in self.peers:
                    start = int(plan.peer_ghost_start[r][me])
                    base = int(hdl.buffer_ptrs[r]) + start * row_bytes
                    addr.append(base + row_bytes * torch.arange(int(plan.send_counts[r]), dtype=torch.int64))
                sym_arrays.append(sym[:a.shape[0]])
                handles.append(hdl)
                dst_addr.append(torch.cat(addr).to(dev) if addr else None)
            self.sets.append((sym_arrays, handles, dst_addr))
        self.parity = 0
        self.barriers_per_step = 1 if double_buffer else 2

    @property
    def arrays(self):
        """The buffer set the next force evaluation reads."""
        return self.sets[self.parity][0]

    def bytes_per_step(self):
        n = int(sum(self.plan.recv_counts[r] for r in range(self.plan.world) if r != self.plan.rank))
        return sum(n * a.shape[1] * a.element_size() for a in self.arrays)

    def advance(self):
        """Switch to the other buffer set (call before filling this step's local particles)."""
        if len(self.sets) == 2:
            self.parity ^= 1
        return self.arrays

    def __call__(self):
        """Enqueue push + barrier for the current buffer set on the current stream."""
        from . import kernels

        arrays, handles, dst_addr = self.sets[self.parity]
        h = handles[0]
        if len(self.sets) == 1:
            h.barrier(channel=0)  # single set: the readers of the ghosts must be done first
        if self.peers:
            for a, dst in zip(arrays, dst_addr):
                kernels.push_rows(a, self._send_cat, dst)
        h.barrier(channel=1 + self.parity)


class SliceScheduler:
    """One rank's slice of a workload: local+ghost State, remapped list, potentials, exchange."""

    def __init__(self, plan, state, nlist, pots, exchange_arrays, group=None, transport="nccl"):
        self.plan = plan
        self.state = state
        self.nlist = nlist
        self.pots = pots
        self.exchange_arrays = exchange_arrays
        self.transport = transport
        self.halo = HaloExchange(plan, group)
        self.peer_halo = None
        if transport == "peer":
            # move the exchanged arrays into symmetric memory; the State keeps views of them
            try:
                self.peer_halo = PeerHalo(plan, exchange_arrays, group)
            except Exception as exc:  # no peer access / symmetric memory on this system
                import warnings

                warnings.warn("peer-memory halo unavailable (%s: %s); using the NCCL transport"
                              % (type(exc).__name__, exc))
            # every rank must take the same transport
            ok = torch.tensor([1 if self.peer_halo is not None else 0], device=state.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self.peer_halo = None
                self.transport = transport = "nccl"
        if self.peer_halo is not None:
            self._exchanged_names = []
            for old in exchange_arrays:
                for name in ("pos", "vel", "orientation"):
                    if getattr(state, name) is old:
                        self._exchanged_names.append(name)
            self._bind_buffer_set()
        elif transport != "nccl":
            raise ValueError("transport must be 'nccl' or 'peer'")
        self.n_local = plan.n_local
        # high priority: the few CTAs of the pack / NCCL send-recv kernels must be scheduled
        # between the thousands of CTAs of the interior-row kernel, not after them
        self.comm_stream = torch.cuda.Stream(device=state.device, priority=-1)
        self._halo_done = torch.cuda.Event()
        self._packed = torch.cuda.Event()
        self._step_done = torch.cuda.Event()
        self._step_done.record()
        if self.peer_halo is not None:
            self.launches_per_step = len(pots) + len(self.exchange_arrays)
        else:
            self.launches_per_step = len(pots) * ((1 if plan.interior_rows.numel() else 0)
                                                  + (1 if plan.boundary_rows.numel() else 0))
        self._host = None

    def _bind_buffer_set(self):
        """Point the State (what the force kernels read) at the current symmetric buffer set."""
        arrays = self.peer_halo.arrays
        for name, arr in zip(self._exchanged_names, arrays):
            setattr(self.state, name, arr)
        self.exchange_arrays = list(arrays)

    # ---- construction from a synthetic workload ------------------------------------------
    @classmethod
    def from_workload(cls, wl, rank, world, device, dtype=np.float32, buffer=0.4, group=None,
                      transport="nccl", balance="neighbors"):
        """Every rank generates the same global workload (seeded), builds the rows of its slice
        on its GPU, derives the plan and keeps only local + ghost particles.

        ``balance``: "neighbors" (default) cuts the sorted particle array where the running sum
        of n_neigh reaches equal shares (rows of C3 are skewed, 1,860 entries for
        a colloid against ~100 for a solvent particle) -- every rank counts the rows of the whole
        system once at setup for that; "count" cuts at equal particle counts."""
        from . import nlist as aznlist
        from .state import State

        g = wl.make_state(dtype=dtype, device=device)  # global arrays, setup only
        cell = aznlist.Cell(buffer=buffer)
        probe = wl.make_potentials(cell)  # registers the cutoffs with the list
        if balance == "neighbors" and world > 1:
            cell.build(g)  # all rows: the same list on every rank
            weights = cell.n_neigh[:wl.N].cpu().numpy().astype(np.float64)
            bounds = partition_bounds_weighted(weights, world)
            lo, hi = int(bounds[rank]), int(bounds[rank + 1])
            start = int(cell.head_list[lo].item())
            end = int(cell.head_list[hi].item()) if hi < wl.N else int(cell.size)
            n_neigh = cell.n_neigh[lo:hi].clone()
            head_list = (cell.head_list[lo:hi] - start).clone()
            nlist_rows = cell.nlist[start:end].clone()
            n_max = cell.n_max
        elif balance in ("count", "neighbors"):
            bounds = partition_bounds(wl.N, world)
            lo, hi = int(bounds[rank]), int(bounds[rank + 1])
            cell.build(g, rows=(lo, hi))
            n_neigh, head_list, nlist_rows, n_max = cell.n_neigh, cell.head_list, cell.nlist, cell.n_max
        else:
            raise ValueError("balance must be 'neighbors' or 'count'")
        plan = SlicePlan(rank, world, bounds, n_neigh, head_list, nlist_rows)
        plan.negotiate(group)
        ids = torch.cat([torch.arange(lo, hi, device=g.pos.device), plan.ghost_ids])
        state = State.__new__(State)
        state.box, state.types, state.dtype, state.device = g.box, g.types, g.dtype, g.device
        state.seed, state.timestep, state.dt = g.seed, g.timestep, g.dt
        state.N, state.n_ghost = plan.n_local, plan.n_ghost
        state.pos = g.pos[ids].contiguous()
        state.vel = g.vel[ids].contiguous()
        state.orientation = g.orientation[ids].contiguous()
        state.tag = g.tag[ids].contiguous()
        local_list = aznlist.NeighborList.from_arrays(plan.n_neigh, plan.nlist_local,
                                                      plan.head_list, device=device, buffer=buffer)
        del g, cell, probe, n_neigh, head_list, nlist_rows
        torch.cuda.empty_cache()
        pots = wl.make_potentials(local_list)
        for p in pots:
            p.attach(state)
        arrays = [state.pos]
        names = {type(p).__name__ for p in pots}
        if "DPDGeneralWeight" in names:
            arrays.append(state.vel)
        if "TwoPatchMorse" in names:
            arrays.append(state.orientation)
        return cls(plan, state, local_list, pots, arrays, group, transport)

    # ---- per step ------------------------------------------------------------------------
    def exchange_bytes_per_step(self):
        if self.peer_halo is not None:
            return self.peer_halo.bytes_per_step()
        return self.halo.bytes_per_step(self.exchange_arrays)

    def mean_row_length(self):
        return float(self.nlist.n_neigh.double().mean().item())

    def step(self, compute_virial=False):
        p = self.plan
        if self.peer_halo is not None:
            # next buffer set, push over NVLink, one barrier, then every row in one launch per
            # potential. (In an MD loop the integrator writes the new local positions into the
            # set that advance() returns; here the particles do not move, both sets hold them.)
            self.peer_halo.advance()
            self._bind_buffer_set()
            self.peer_halo()
            for pot in self.pots:
                pot.compute(compute_virial=compute_virial)
            return
        cur = torch.cuda.current_stream()
        # the exchange may overwrite ghosts only after the previous step's boundary rows are done
        self.comm_stream.wait_event(self._step_done)
        # pack on the compute stream (microseconds), then the interior rows, and only the NCCL
        # send/recv on the communication stream: measured on B200, NCCL's two CTAs slip in between
        # the CTAs of the interior kernel, while a bandwidth kernel on a second stream does not
        # (the interior kernel holds every register file), so the pack must not sit there
        cur.wait_event(self._step_done)
        self.halo.pack(self.exchange_arrays)
        self._packed.record()
        if p.interior_rows.numel():
            for pot in self.pots:
                pot.compute(compute_virial=compute_virial, row_ids=p.interior_rows)
        self.comm_stream.wait_event(self._packed)
        with torch.cuda.stream(self.comm_stream):
            self.halo.exchange(self.exchange_arrays)
            self._halo_done.record()
        cur.wait_event(self._halo_done)
        if p.boundary_rows.numel():
            for pot in self.pots:
                pot.compute(compute_virial=compute_virial, row_ids=p.boundary_rows)
        self._step_done.record()

    def tune(self, compute_virial=False):
        """Autotune each potential's launch shape on the interior rows."""
        out = []
        if self.peer_halo is not None:
            self.peer_halo()
        else:
            self.halo(self.exchange_arrays)
        torch.cuda.synchronize()
        for pot in self.pots:
            was_frozen = pot.nlist._frozen
            pot.nlist.freeze()
            rows = None
            if self.peer_halo is None and self.plan.interior_rows.numel():
                rows = self.plan.interior_rows
            args = pot._args(None, compute_virial, rows)
            from . import kernels

            b, t, ms = kernels.autotune(pot._family, pot._evaluator, pot._bits, args,
                                        pot._d_params.data_ptr())
            pot.kernel_parameters = (b, t)
            out.append((b, t, ms))
            pot.nlist._frozen = was_frozen
        return out

    def time_kernels(self, steps, compute_virial=False):
        """ms per step of the force kernels alone (no exchange), CUDA events."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            if self.peer_halo is not None:
                for pot in self.pots:
                    pot.compute(compute_virial=compute_virial)
                continue
            for rows in (self.plan.interior_rows, self.plan.boundary_rows):
                if rows.numel():
                    for pot in self.pots:
                        pot.compute(compute_virial=compute_virial, row_ids=rows)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    # ---- correctness of the sliced evaluation ------------------------------------------------
    def verify_against_single_domain(self, wl, m=4096, compute_virial=False, buffer=0.4):
        """Compare rows of this rank's slice -- the first ``m`` (they reference ghosts delivered by
        the halo exchange) and ``m`` from the middle -- with an evaluation of the same rows on a
        single-domain copy of the whole system held on this GPU (same kernels, same launch
        shape, global indices instead of local + ghost ones). Called after steps have run, so
        it checks what the exchange actually delivered. Returns a dict with the largest
        per-particle force (+ torque) error relative to the rms force and the bit-identity flag
        (DPD: the deferred-accept queue makes the fp32 summation order depend on the row
        partition, so only the relative error applies there)."""
        from . import nlist as aznlist

        lo, n = self.plan.lo, self.n_local
        g = wl.make_state(dtype=self.state.dtype, device=self.state.device)
        cell = aznlist.Cell(buffer=buffer)
        pots_g = wl.make_potentials(cell)
        windows = [(0, min(m, n))]
        if n > 2 * m:
            windows.append((n // 2, n // 2 + m))
        worst, identical, rows = 0.0, True, 0
        for pot, pot_g in zip(self.pots, pots_g):
            pot_g.attach(g)
            pot_g.kernel_parameters = pot.kernel_parameters
            pot.compute(compute_virial=compute_virial)
            for a, b in windows:
                pot_g.compute(compute_virial=compute_virial, rows=(lo + a, lo + b))
                pairs = [(pot._force[a:b], pot_g._force[lo + a:lo + b])]
                if pot.is_anisotropic:
                    pairs.append((pot._torque[a:b], pot_g._torque[lo + a:lo + b]))
                if compute_virial:
                    pairs.append((pot._virial[:, a:b], pot_g._virial[:, lo + a:lo + b]))
                for x, y in pairs:
                    identical = identical and bool(torch.equal(x, y))
                    scale = float(y.double().pow(2).mean().sqrt().item()) or 1.0
                    worst = max(worst, float((x.double() - y.double()).abs().max().item()) / scale)
                rows += b - a
        del g, cell, pots_g
        torch.cuda.empty_cache()
        return dict(rows=rows, max_rel_diff=worst, bit_identical=identical)

    # ---- end to end with host buffers -------------------------------------------------------
    def _host_buffers(self, compute_virial):
        if self._host is None:
            st = self.state
            self._host = dict(
                pos=torch.empty((self.n_local, 4), dtype=st.pos.dtype).pin_memory(),
                force=torch.empty_like(self.pots[0]._force, device="cpu").pin_memory(),
                virial=torch.empty_like(self.pots[0]._virial, device="cpu").pin_memory(),
                torque=torch.empty_like(self.pots[0]._torque, device="cpu").pin_memory())
            self._host["pos"].copy_(st.pos[:self.n_local])
        return self._host

    def e2e_bytes(self, compute_virial):
        h = self._host_buffers(compute_virial)
        h2d = h["pos"].numel() * h["pos"].element_size()
        d2h = h["force"].numel() * h["force"].element_size() * len(self.pots)
        if compute_virial:
            d2h += h["virial"].numel() * h["virial"].element_size() * len(self.pots)
        d2h += sum(h["torque"].numel() * h["torque"].element_size() for p in self.pots if p.is_anisotropic)
        return h2d, d2h

    def e2e_step(self, compute_virial=False):
        """One step through host buffers: this rank's positions from pinned host memory, the halo
        exchange, and the per-particle results delivered to pinned host memory in row chunks
        whose device-to-host copies overlap the evaluation of the next chunk
        (``Pair.compute_to_host``)."""
        h = self._host_buffers(compute_virial)
        if self.peer_halo is not None:
            self.peer_halo.advance()
            self._bind_buffer_set()
        self.state.pos[:self.n_local].copy_(h["pos"], non_blocking=True)
        if self.peer_halo is not None:
            self.peer_halo()
        else:
            self._step_done.record()
            self.halo(self.exchange_arrays)
        for pot in self.pots:
            pot.compute_to_host(h["force"], h["virial"] if compute_virial else None,
                                h["torque"] if pot.is_anisotropic else None)
