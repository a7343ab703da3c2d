"""Neighbour lists in HOOMD's ``NeighborListGPU`` layout.

``Cell`` mirrors ``hoomd.md.nlist.Cell(buffer, ...)`` (the list every reference test uses,
reference src/pytest/test_pair.py:337): a full list with r_list = r_cut(type pair) + buffer,
built on the GPU by ``azp_nlist_*`` (csrc/nlist_kernels.cu) and rebuilt when any particle has
moved more than buffer/2 since the last build. ``NeighborList.from_arrays`` wraps arrays built
elsewhere (e.g. by HOOMD itself) without copying semantics changes.
"""

import ctypes

import numpy as np
import torch

from . import _lib


class NeighborList:
    """Holder of ``n_neigh`` (u32[N]), ``nlist`` (u32[size]) and ``head_list`` (u64[N])."""

    storage_mode = "full"

    def __init__(self, buffer=0.4, exclusions=(), rebuild_check_delay=1, check_dist=True,
                 default_r_cut=0.0):
        self.buffer = float(buffer)
        self.exclusions = tuple(exclusions)
        self.rebuild_check_delay = int(rebuild_check_delay)
        self.check_dist = bool(check_dist)
        self.default_r_cut = float(default_r_cut)
        self._consumers = []
        self.n_neigh = None
        self.nlist = None
        self.head_list = None
        self.size = 0
        self.n_max = 0  # largest row capacity (HOOMD's n_max)
        self.num_builds = 0
        self._pos_at_build = None
        self._moved_flag = None
        self._external = False
        self._frozen = False      # freeze(): benchmarks / graph capture keep the current list
        self._built_for = None    # (box, r_list matrix) of the last build

    # ---- consumers: the list must cover the largest r_cut of every attached potential ------
    def _add_consumer(self, force):
        if force not in self._consumers:
            self._consumers.append(force)
            if not self._external:
                self.n_neigh = None  # r_cut may have grown: rebuild at the next compute

    def r_cut_matrix(self, state):
        nt = state.ntypes
        rc = np.full((nt, nt), self.default_r_cut, dtype=np.float64)
        for f in self._consumers:
            rc = np.maximum(rc, f._r_cut_matrix(state))
        return rc

    @classmethod
    def from_arrays(cls, n_neigh, nlist, head_list, device="cuda:0", buffer=0.0):
        """Adopt an existing HOOMD-layout list (numpy or torch arrays)."""
        self = cls(buffer=buffer)
        dev = torch.device(device)

        def conv(a, np_dtype, view):
            if isinstance(a, torch.Tensor):
                return a.to(dev).contiguous()
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np_dtype).view(view)).to(dev)

        self.n_neigh = conv(n_neigh, np.uint32, np.int32)
        self.nlist = conv(nlist, np.uint32, np.int32)
        self.head_list = conv(head_list, np.uint64, np.int64)
        self.size = int(self.nlist.numel())
        self.n_max = int(self.n_neigh.max().item()) if self.n_neigh.numel() else 0
        self._external = True
        return self

    def freeze(self):
        """Keep the current list whatever the particles do (benchmarks that time the force path
        alone, CUDA-graph capture of a step, launch tuning). Private to this package's drivers:
        HOOMD has no such switch -- its ``check_dist=False`` means the opposite (rebuild at every
        check). Undo with :meth:`thaw`."""
        self._frozen = True
        return self

    def thaw(self):
        self._frozen = False
        return self

    def _signature(self, state):
        b = state.box
        # the consumers bump _tables_version on every params / r_cut / r_on assignment
        return (tuple(b.L), b.xy, b.xz, b.yz, tuple(bool(p) for p in b.periodic), self.buffer,
                tuple(getattr(f, "_tables_version", 0) for f in self._consumers))

    def compute(self, state):
        """Rebuild when needed (HOOMD ``NeighborList::compute``): never for an adopted list; always
        when there is none yet or when the box or an ``r_cut`` changed since the last build;
        otherwise, once ``rebuild_check_delay`` steps have passed since the last build, when some
        particle has moved more than half the buffer -- or unconditionally at every check when
        ``check_dist`` is False (HOOMD's meaning of that flag)."""
        if self._external:
            return False
        if self.n_neigh is not None and self._frozen:
            return False
        if self.n_neigh is None or self._built_for != self._signature(state):
            self.build(state)
            self._build_step = state.timestep
            self._built_for = self._signature(state)
            return True
        if self.steps_until_check(state) > 0:
            return False
        if self._needs_rebuild(state):
            self.build(state)
            self._build_step = state.timestep
            return True
        return False

    def steps_until_check(self, state):
        """Time steps from now during which no displacement check is due (0 = check now)."""
        if self.rebuild_check_delay <= 1:
            return 0  # the default: every compute() checks (also repeated calls at one time step)
        since = state.timestep - getattr(self, "_build_step", state.timestep - 10 ** 9)
        return max(0, self.rebuild_check_delay - since)

    def build(self, state):
        raise NotImplementedError

    def _needs_rebuild(self, state):
        if not self.check_dist or self._pos_at_build is None:
            return True  # check_dist=False: rebuild at every check (HOOMD semantics)
        if self._pos_at_build.shape != state.pos.shape:
            return True
        # device-side check (azp_nlist_moved): one kernel and one 4-byte read-back
        if self._moved_flag is None or self._moved_flag.device != state.pos.device:
            self._moved_flag = torch.zeros(1, dtype=torch.int32, device=state.pos.device)
        self._moved_flag.zero_()
        box = state.box.to_c()
        with torch.cuda.device(state.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            fn = getattr(_lib.lib, "azp_nlist_moved_f%d" % (8 * state.dtype.itemsize))
            _lib.check(fn(state.pos.data_ptr(), self._pos_at_build.data_ptr(), ctypes.byref(box),
                          0.5 * self.buffer, state.pos.shape[0], self._moved_flag.data_ptr(), stream),
                       "nlist moved")
        return bool(self._moved_flag.item())

    def to_numpy(self):
        return (self.n_neigh.cpu().numpy().view(np.uint32),
                self.nlist.cpu().numpy().view(np.uint32),
                self.head_list.cpu().numpy().view(np.uint64))


class Cell(NeighborList):
    """Cell-list neighbour search on the GPU; ctor mirrors ``hoomd.md.nlist.Cell``."""

    def __init__(self, buffer, exclusions=("bond",), rebuild_check_delay=1, check_dist=True,
                 deterministic=False, mesh=None, default_r_cut=0.0, row_align=8, threads_per_row=0):
        super().__init__(buffer, exclusions, rebuild_check_delay, check_dist, default_r_cut)
        self.deterministic = deterministic
        self.row_align = int(row_align)
        self.threads_per_row = int(threads_per_row)  # lanes per row of the builder (0 = library choice)
        self.reuse_capacity = True  # skip the count pass while the previous row capacities suffice

    def build(self, state, rows=None):
        """Build the list for all particles of ``state`` or, with ``rows=(lo, hi)``, only for the
        rows of particles [lo, hi) (neighbours are still searched among all particles and stored
        as global indices; n_neigh / head_list are then indexed by row - lo)."""
        if not state.pos.is_cuda:
            raise _lib.AzpError("nlist.Cell builds on the GPU only (no CPU fallback)")
        bits = 8 * state.dtype.itemsize
        sfx = "_f%d" % bits
        nt = state.ntypes
        r_list = self.r_cut_matrix(state) + self.buffer
        r_list[self.r_cut_matrix(state) <= 0] = 0.0
        r_max = float(r_list.max())
        if not r_max > 0:
            raise ValueError("neighbour list has no positive r_cut")
        for d, width in enumerate(state.box.nearest_plane_distance()):
            if state.box.periodic[d] and width < 2.0 * r_max:
                raise ValueError("box too small for r_cut + buffer (minimum image)")
        dev = state.device
        n_total = state.pos.shape[0]
        lo, hi = (0, n_total) if rows is None else (int(rows[0]), int(rows[1]))
        n_rows = hi - lo
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            rl = r_list.astype(state.dtype)
            rlsq = torch.from_numpy((rl * rl).reshape(-1).copy()).to(dev)
            a = _lib.AzpNlistArgs()
            a.d_pos = state.pos.data_ptr()
            a.N = n_total
            a.ntypes = nt
            a.box = state.box.to_c()
            a.d_rlistsq = rlsq.data_ptr()
            a.r_list_max = r_max
            dim = (ctypes.c_uint32 * 3)()
            _lib.check(_lib.lib.azp_nlist_cell_dim(ctypes.byref(a.box), r_max, dim), "cell_dim")
            ncells = int(dim[0]) * int(dim[1]) * int(dim[2])
            for d in range(3):
                a.cell_dim[d] = dim[d]
            cell_of = torch.empty(n_total, dtype=torch.int32, device=dev)
            cell_start = torch.empty(ncells + 1, dtype=torch.int32, device=dev)
            cell_order = torch.empty(n_total, dtype=torch.int32, device=dev)
            cell_pos = torch.empty_like(state.pos)
            if (self._pos_at_build is None or self._pos_at_build.shape != state.pos.shape
                    or self._pos_at_build.dtype != state.pos.dtype or self._pos_at_build.device != dev):
                self._pos_at_build = torch.empty_like(state.pos)
            # persistent buffers where the shape allows: consumers that recorded device
            # addresses (a CUDA graph of the MD step) stay valid across rebuilds
            if (self.n_neigh is not None and not self._external and self.n_neigh.numel() == n_rows
                    and self.n_neigh.device == dev):
                n_neigh = self.n_neigh
            else:
                n_neigh = torch.empty(n_rows, dtype=torch.int32, device=dev)
            a.row_offset = lo
            a.n_rows = n_rows
            a.d_cell_of = cell_of.data_ptr()
            a.d_cell_start = cell_start.data_ptr()
            a.d_cell_order = cell_order.data_ptr()
            a.d_cell_pos = cell_pos.data_ptr()
            a.d_pos_at_build = self._pos_at_build.data_ptr()  # bin copies the positions there
            a.threads_per_row = self.threads_per_row
            a.d_n_neigh = n_neigh.data_ptr()
            _lib.check(getattr(_lib.lib, "azp_nlist_bin" + sfx)(ctypes.byref(a), stream), "nlist bin")
            # Capacity reuse (what HOOMD's NeighborList does between builds): when the previous
            # build had the same rows, fill straight into its row capacities -- the fill counts as
            # it goes -- and fall back to count + fill only if some row overflowed.
            prev = getattr(self, "_reuse", None)
            reused = False
            if prev is not None and prev["key"] == (n_total, lo, hi, bits) and self.reuse_capacity:
                head, cap32, size = prev["head"], prev["cap32"], prev["size"]
                nlist = self.nlist if (self.nlist is not None and self.nlist.numel() == max(size, 1)) \
                    else torch.zeros(max(size, 1), dtype=torch.int32, device=dev)
                prev["overflow"].zero_()
                a.d_head_list = head.data_ptr()
                a.d_nlist = nlist.data_ptr()
                a.d_capacity = cap32.data_ptr()
                a.d_overflow = prev["overflow"].data_ptr()
                _lib.check(getattr(_lib.lib, "azp_nlist_fill" + sfx)(ctypes.byref(a), stream), "nlist fill")
                reused = int(prev["overflow"].item()) == 0
                a.d_capacity = None
                a.d_overflow = None
            if not reused:
                _lib.check(getattr(_lib.lib, "azp_nlist_count" + sfx)(ctypes.byref(a), stream), "nlist count")
                # head_list = prefix sum of Nmax[type] (HOOMD's row capacity rule)
                typeid = particle_typeid(state.pos)[lo:hi]
                cap = torch.zeros(n_rows, dtype=torch.int64, device=dev)
                for t in range(nt):
                    sel = typeid == t
                    if bool(sel.any()):
                        m = int(n_neigh[sel].max())
                        m = (m + self.row_align - 1) // self.row_align * self.row_align
                        cap[sel] = m
                head = torch.cumsum(cap, 0) - cap
                size = int(cap.sum())
                self.n_max = int(cap.max().item()) if n_rows else 0
                nlist = torch.zeros(max(size, 1), dtype=torch.int32, device=dev)
                a.d_head_list = head.data_ptr()
                a.d_nlist = nlist.data_ptr()
                _lib.check(getattr(_lib.lib, "azp_nlist_fill" + sfx)(ctypes.byref(a), stream), "nlist fill")
                self._reuse = dict(key=(n_total, lo, hi, bits), head=head, cap32=cap.to(torch.int32),
                                   size=size, overflow=torch.zeros(1, dtype=torch.int32, device=dev))
            self.num_reused = getattr(self, "num_reused", 0) + int(reused)
        # rows of ghosts are built too but only the first N rows are consumed
        self.n_neigh, self.nlist, self.head_list, self.size = n_neigh, nlist, head, size
        self.num_builds += 1


def particle_typeid(pos):
    """Type ids bit-cast in pos[:, 3] (torch tensor, fp32 or fp64)."""
    if pos.dtype == torch.float32:
        return pos[:, 3].contiguous().view(torch.int32).to(torch.int64)
    return pos[:, 3].contiguous().view(torch.int64) & 0xFFFFFFFF
