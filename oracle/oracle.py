"""TEST INFRASTRUCTURE (oracle) -- ctypes front end of the CPU checker libraries.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module. The product package ``azplugins_b200`` never
does; its hot path is the CUDA extension and fails loudly without it.

Two kinds of library export the same C symbols (``oracle/oracle_main.cc``):

* ``kind="port"``  -- ``oracle/liboracle_port_f{32,64}.so``: our restatement of the evaluators
  (``port_evaluators.h``) under our restatement of HOOMD's CPU loops (``driver_loops.h``).
* ``kind="ref"``   -- ``oracle/_ref/liboracle_ref_f{32,64}.so``: the reference's own evaluator
  headers compiled in place from ``/root/reference/src`` under the same loops. Built only where
  ``/root/reference`` exists; the prebuilt files travel to the GPU box.

``kind="best"`` picks ``ref`` when present, else ``port``.
"""

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

EVALUATORS = {
    "PerturbedLennardJones": 0,
    "ExpandedYukawa": 1,
    "Colloid": 2,
    "Hertz": 3,
    "DPDGeneralWeight": 4,
    "TwoPatchMorse": 5,
}
# order of the double fields handed to *_pack_params (same as include/azp_b200.h)
PARAM_FIELDS = {
    "PerturbedLennardJones": ("epsilon", "sigma", "attraction_scale_factor"),
    "ExpandedYukawa": ("epsilon", "kappa", "delta"),
    "Colloid": ("A", "a_1", "a_2", "sigma"),
    "Hertz": ("epsilon",),
    "DPDGeneralWeight": ("A", "gamma", "s"),
    "TwoPatchMorse": ("M_d", "M_r", "r_eq", "omega", "alpha", "repulsion"),
}
SHIFT_MODES = {"none": 0, "shift": 1, "xplor": 2}


class _Args(ctypes.Structure):
    _fields_ = [
        ("N", ctypes.c_uint32),
        ("ntypes", ctypes.c_uint32),
        ("pos", ctypes.c_void_p),
        ("n_neigh", ctypes.c_void_p),
        ("nlist", ctypes.c_void_p),
        ("head_list", ctypes.c_void_p),
        ("L", ctypes.c_double * 3),
        ("tilt", ctypes.c_double * 3),
        ("periodic", ctypes.c_int32 * 3),
        ("shift_mode", ctypes.c_int32),
        ("compute_virial", ctypes.c_int32),
        ("half_list", ctypes.c_int32),
        ("rint_image", ctypes.c_int32),
        ("nthreads", ctypes.c_int32),
        ("rcutsq", ctypes.c_void_p),
        ("ronsq", ctypes.c_void_p),
        ("force", ctypes.c_void_p),
        ("virial", ctypes.c_void_p),
        ("virial_pitch", ctypes.c_uint64),
        ("vel", ctypes.c_void_p),
        ("tag", ctypes.c_void_p),
        ("seed", ctypes.c_uint32),
        ("timestep", ctypes.c_uint64),
        ("deltaT", ctypes.c_double),
        ("T", ctypes.c_double),
        ("orientation", ctypes.c_void_p),
        ("torque", ctypes.c_void_p),
    ]


def build(ref=None, quiet=True):
    """Compile the oracle libraries (``make -C oracle``). ``ref=None`` builds ``_ref`` only when
    /root/reference is present."""
    targets = ["port"]
    if ref is None:
        ref = os.path.isdir("/root/reference/src")
    if ref:
        targets.append("ref")
    cmd = ["make", "-C", HERE, "-j4"] + targets
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL if quiet else None)


def _path(kind, bits):
    if kind == "port":
        return os.path.join(HERE, "liboracle_port_f%d.so" % bits)
    return os.path.join(HERE, "_ref", "liboracle_ref_f%d.so" % bits)


def available(kind, bits=32):
    return os.path.exists(_path(kind, bits))


_CACHE = {}


def load(kind="best", dtype=np.float32):
    """Return an :class:`Oracle` for ``kind`` in {"port", "ref", "best"} and a numpy float dtype."""
    dtype = np.dtype(dtype)
    bits = 8 * dtype.itemsize
    if kind == "best":
        kind = "ref" if available("ref", bits) else "port"
    key = (kind, bits)
    if key not in _CACHE:
        path = _path(kind, bits)
        if not os.path.exists(path):
            if kind == "port":
                build(ref=False)
            else:
                raise FileNotFoundError(path + " (build with `make -C oracle ref`)")
        _CACHE[key] = Oracle(path, kind, dtype)
    return _CACHE[key]


class Oracle:
    def __init__(self, path, kind, dtype):
        self.kind = kind
        self.dtype = np.dtype(dtype)
        self.path = path
        self.lib = ctypes.CDLL(path)
        lib = self.lib
        assert lib.oracle_scalar_size() == self.dtype.itemsize
        assert bool(lib.oracle_is_reference()) == (kind == "ref")
        for name in ("oracle_pair_forces", "oracle_dpd_forces", "oracle_aniso_forces"):
            getattr(lib, name).argtypes = [ctypes.c_int, ctypes.POINTER(_Args), ctypes.c_void_p]
            getattr(lib, name).restype = ctypes.c_int
        lib.oracle_eval_pair.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_double,
                                         ctypes.c_double, ctypes.c_int, ctypes.c_void_p]
        lib.oracle_eval_dpd_thermo.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double,
                                               ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                               ctypes.c_uint64, ctypes.c_double, ctypes.c_double,
                                               ctypes.c_double, ctypes.c_void_p]
        lib.oracle_eval_aniso.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_double, ctypes.c_int,
                                          ctypes.c_void_p]
        lib.oracle_dpd_alpha.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                         ctypes.c_uint64]
        lib.oracle_dpd_alpha.restype = ctypes.c_double
        lib.oracle_philox.argtypes = [ctypes.c_void_p] * 3
        lib.oracle_philox.restype = None
        lib.oracle_min_image.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_void_p]
        lib.oracle_min_image.restype = None
        lib.oracle_wall.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                    ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_uint64]
        lib.oracle_wall.restype = ctypes.c_int
        lib.oracle_barrier.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_uint32, ctypes.c_void_p,
                                       ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int]
        lib.oracle_barrier.restype = ctypes.c_int
        lib.oracle_nlist.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int, ctypes.c_uint32]
        if kind == "port":
            lib.oracle_pack_params.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        else:
            self.pylib = ctypes.PyDLL(path)  # GIL held: these entry points build Python objects
            self.pylib.oracle_ref_pack_params.argtypes = [ctypes.c_int, ctypes.py_object,
                                                          ctypes.c_void_p]
            self.pylib.oracle_ref_pack_params.restype = ctypes.c_int
            self.pylib.oracle_ref_unpack_params.argtypes = [ctypes.c_int, ctypes.c_void_p]
            self.pylib.oracle_ref_unpack_params.restype = ctypes.py_object
            self.pylib.oracle_ref_name.argtypes = [ctypes.c_int]
            self.pylib.oracle_ref_name.restype = ctypes.py_object

    # ---- parameters ----------------------------------------------------------------------
    def max_threads(self):
        return int(self.lib.oracle_max_threads())

    def param_size(self, evaluator):
        return int(self.lib.oracle_param_size(EVALUATORS[evaluator]))

    def pack_params(self, evaluator, params):
        """``param_type`` bytes for one parameter dict (``np.uint8[param_size]``).

        ``ref`` runs the reference's own ``param_type(pybind11::dict)`` constructor; ``port`` the
        restated one."""
        ev = EVALUATORS[evaluator]
        out = np.zeros(self.param_size(evaluator), dtype=np.uint8)
        if self.kind == "ref":
            rc = self.pylib.oracle_ref_pack_params(ev, dict(params), out.ctypes.data)
        else:
            fields = np.array([float(params[k]) for k in PARAM_FIELDS[evaluator]],
                              dtype=np.float64)
            rc = self.lib.oracle_pack_params(ev, fields.ctypes.data, out.ctypes.data)
        if rc != 0:
            raise RuntimeError("pack_params failed")
        return out

    def unpack_params(self, evaluator, raw):
        assert self.kind == "ref"
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        return self.pylib.oracle_ref_unpack_params(EVALUATORS[evaluator], raw.ctypes.data)

    def name(self, evaluator):
        assert self.kind == "ref"
        return self.pylib.oracle_ref_name(EVALUATORS[evaluator])

    def pack_table(self, evaluator, ntypes, pair_params):
        """``param_type[ntypes^2]`` bytes from ``{(ti, tj): dict}`` stored symmetrically at
        Index2D(i, j) = j * ntypes + i (reference call stack 3.1 step 4)."""
        sz = self.param_size(evaluator)
        table = np.zeros((ntypes * ntypes, sz), dtype=np.uint8)
        for (ti, tj), p in pair_params.items():
            raw = self.pack_params(evaluator, p)
            table[tj * ntypes + ti] = raw
            table[ti * ntypes + tj] = raw
        return table

    # ---- single-pair probes --------------------------------------------------------------
    def eval_pair(self, evaluator, params, rsq, rcutsq, shift=False):
        raw = self.pack_params(evaluator, params)
        out = np.zeros(3)
        rc = self.lib.oracle_eval_pair(EVALUATORS[evaluator], raw.ctypes.data, rsq, rcutsq,
                                       int(shift), out.ctypes.data)
        assert rc == 0
        return bool(out[0]), out[1], out[2]

    def eval_dpd_thermo(self, params, rsq, rcutsq, seed, tag_i, tag_j, timestep, dt, rdotv, T):
        raw = self.pack_params("DPDGeneralWeight", params)
        out = np.zeros(4)
        self.lib.oracle_eval_dpd_thermo(raw.ctypes.data, rsq, rcutsq, seed, tag_i, tag_j,
                                        timestep, dt, rdotv, T, out.ctypes.data)
        return bool(out[0]), out[1], out[2], out[3]

    def eval_aniso(self, params, dr, qi, qj, rcutsq, shift=False):
        raw = self.pack_params("TwoPatchMorse", params)
        dr = np.asarray(dr, dtype=np.float64)
        qi = np.asarray(qi, dtype=np.float64)
        qj = np.asarray(qj, dtype=np.float64)
        out = np.zeros(11)
        self.lib.oracle_eval_aniso(raw.ctypes.data, dr.ctypes.data, qi.ctypes.data,
                                   qj.ctypes.data, rcutsq, int(shift), out.ctypes.data)
        return bool(out[0]), out[1:4].copy(), out[4], out[5:8].copy(), out[8:11].copy()

    def dpd_alpha(self, seed, tag_i, tag_j, timestep):
        return float(self.lib.oracle_dpd_alpha(seed, tag_i, tag_j, timestep))

    def philox(self, ctr, key):
        c = np.asarray(ctr, dtype=np.uint32)
        k = np.asarray(key, dtype=np.uint32)
        out = np.zeros(4, dtype=np.uint32)
        self.lib.oracle_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
        return out

    def min_image(self, v, L, tilt=(0, 0, 0), periodic=(1, 1, 1), rint=False):
        v = np.array(v, dtype=np.float64)
        L = np.asarray(L, dtype=np.float64)
        t = np.asarray(tilt, dtype=np.float64)
        p = np.asarray(periodic, dtype=np.int32)
        self.lib.oracle_min_image(L.ctypes.data, t.ctypes.data, p.ctypes.data, int(rint),
                                  v.ctypes.data)
        return v

    # ---- neighbour list ------------------------------------------------------------------
    def build_nlist(self, pos, L, r_list, ntypes=1, tilt=(0, 0, 0), periodic=(1, 1, 1),
                    half=False, nthreads=0, row_align=8, n_rows=0):
        """HOOMD-layout list. ``r_list``: scalar or (ntypes, ntypes) array of r_cut + buffer.
        Returns (n_neigh u32[N], nlist u32[size], head_list u64[N]). Row capacity is the per-type
        maximum rounded up to ``row_align`` (HOOMD: head_list = prefix sum of Nmax[type]).
        ``n_rows`` > 0 builds only rows [0, n_rows) (the others stay empty): bounded CPU samples."""
        pos = np.ascontiguousarray(pos, dtype=self.dtype)
        N = pos.shape[0]
        rl = np.broadcast_to(np.asarray(r_list, dtype=np.float64), (ntypes, ntypes))
        rlsq = np.ascontiguousarray((rl * rl).astype(self.dtype)).reshape(-1)
        Ld = np.asarray(L, dtype=np.float64)
        td = np.asarray(tilt, dtype=np.float64)
        pd = np.asarray(periodic, dtype=np.int32)
        nt = nthreads if nthreads > 0 else self.max_threads()
        n_neigh = np.zeros(N, dtype=np.uint32)
        self.lib.oracle_nlist(0, N, pos.ctypes.data, Ld.ctypes.data, td.ctypes.data,
                              pd.ctypes.data, ntypes, rlsq.ctypes.data, int(half),
                              n_neigh.ctypes.data, None, None, nt, n_rows)
        types = particle_types(pos)
        nmax = np.zeros(ntypes, dtype=np.uint64)
        for t in range(ntypes):
            sel = n_neigh[types == t]
            m = int(sel.max()) if sel.size else 0
            nmax[t] = (m + row_align - 1) // row_align * row_align if row_align > 1 else m
        cap = nmax[types]
        if n_rows:
            cap[n_rows:] = 0
        head = np.zeros(N, dtype=np.uint64)
        if N > 1:
            np.cumsum(cap[:-1], out=head[1:])
        size = int(cap.sum())
        nlist = np.zeros(max(size, 1), dtype=np.uint32)
        self.lib.oracle_nlist(1, N, pos.ctypes.data, Ld.ctypes.data, td.ctypes.data,
                              pd.ctypes.data, ntypes, rlsq.ctypes.data, int(half),
                              n_neigh.ctypes.data, head.ctypes.data, nlist.ctypes.data, nt, n_rows)
        return n_neigh, nlist, head

    # ---- wall potentials (reference src/WallEvaluator*.h under HOOMD's wall loop) -----------
    def wall_forces(self, evaluator, pos, params, spheres=(), cylinders=(), planes=()):
        """evaluator "Colloid" (params rows {c_1, c_2, a, rcutsq, rextrap}) or "LJ93"
        ({sigma_3, A, rcutsq, rextrap}); walls: spheres (r, origin3, inside, open), cylinders
        (r, origin3, unit axis3, inside, open), planes (origin3, unit normal3, open).
        Returns dict(force (N,4), virial (6,N))."""
        pos = np.ascontiguousarray(pos, dtype=self.dtype)
        par = np.ascontiguousarray(params, dtype=self.dtype)
        N = pos.shape[0]
        sph = np.ascontiguousarray(np.asarray(spheres, dtype=np.float64).reshape(-1, 6))
        cyl = np.ascontiguousarray(np.asarray(cylinders, dtype=np.float64).reshape(-1, 9))
        pla = np.ascontiguousarray(np.asarray(planes, dtype=np.float64).reshape(-1, 7))
        force = np.zeros((N, 4), dtype=self.dtype)
        virial = np.zeros((6, max(N, 1)), dtype=self.dtype)
        ev = {"Colloid": 0, "LJ93": 1}[evaluator]
        rc = self.lib.oracle_wall(ev, N, pos.ctypes.data, par.ctypes.data, sph.shape[0],
                                  sph.ctypes.data, cyl.shape[0], cyl.ctypes.data, pla.shape[0],
                                  pla.ctypes.data, force.ctypes.data, virial.ctypes.data,
                                  virial.shape[1])
        if rc != 0:
            raise ValueError("unknown wall evaluator")
        return dict(force=force, virial=virial[:, :N])

    # ---- external harmonic barrier (reference src/HarmonicBarrier.h:149-175) ---------------
    def barrier_forces(self, geometry, location, pos, params, L, tilt=(0, 0, 0),
                       periodic=(1, 1, 1), check_valid=True):
        """geometry: "planar" | "spherical"; params: (ntypes, 2) array of {k, offset}.
        Returns dict(force (N,4), virial (6,N)); raises RuntimeError("Barrier position is
        invalid") like the reference when the barrier lies outside the box."""
        pos = np.ascontiguousarray(pos, dtype=self.dtype)
        par = np.ascontiguousarray(params, dtype=self.dtype).reshape(-1, 2)
        N = pos.shape[0]
        Ld = np.asarray(L, dtype=np.float64)
        td = np.asarray(tilt, dtype=np.float64)
        pd = np.asarray(periodic, dtype=np.int32)
        force = np.zeros((N, 4), dtype=self.dtype)
        virial = np.full((6, max(N, 1)), 7.0, dtype=self.dtype)
        g = {"planar": 0, "spherical": 1}[geometry]
        rc = self.lib.oracle_barrier(g, float(location), N, pos.ctypes.data, par.shape[0],
                                     par.ctypes.data, Ld.ctypes.data, td.ctypes.data,
                                     pd.ctypes.data, force.ctypes.data, virial.ctypes.data,
                                     virial.shape[1], int(check_valid))
        if rc != 0:
            raise RuntimeError("Barrier position is invalid")
        return dict(force=force, virial=virial[:, :N])

    # ---- force loops ---------------------------------------------------------------------
    def _args(self, pos, n_neigh, nlist, head, L, tilt, periodic, ntypes, rcut, ron, mode,
              virial, half, rint_image, nthreads, N=None):
        dt = self.dtype
        keep = []

        def arr(x, dtype):
            a = np.ascontiguousarray(x, dtype=dtype)
            keep.append(a)
            return a

        pos = arr(pos, dt)
        n_rows = int(N if N is not None else pos.shape[0])
        a = _Args()
        a.N = n_rows
        a.ntypes = ntypes
        a.pos = pos.ctypes.data
        a.n_neigh = arr(n_neigh, np.uint32).ctypes.data
        a.nlist = arr(nlist, np.uint32).ctypes.data
        a.head_list = arr(head, np.uint64).ctypes.data
        for d in range(3):
            a.L[d] = float(L[d])
            a.tilt[d] = float(tilt[d])
            a.periodic[d] = int(periodic[d])
        a.shift_mode = SHIFT_MODES[mode] if isinstance(mode, str) else int(mode)
        a.compute_virial = int(bool(virial))
        # half: False (full list), True (half list, third law, one thread), "domains" (full list
        # in, one half-list domain per thread: HOOMD's MPI decomposition; driver_loops.h)
        a.half_list = 2 if half == "domains" else int(bool(half))
        a.rint_image = int(bool(rint_image))
        a.nthreads = nthreads if nthreads > 0 else self.max_threads()
        rc = np.broadcast_to(np.asarray(rcut, dtype=np.float64), (ntypes, ntypes))
        rcsq = arr((rc.astype(dt) * rc.astype(dt)).reshape(-1), dt)
        a.rcutsq = rcsq.ctypes.data
        ro = np.broadcast_to(np.asarray(ron, dtype=np.float64), (ntypes, ntypes))
        rosq = arr((ro.astype(dt) * ro.astype(dt)).reshape(-1), dt)
        a.ronsq = rosq.ctypes.data
        n_out = pos.shape[0] if half is True else n_rows
        force = np.zeros((n_out, 4), dtype=dt)
        pitch = n_out
        vir = np.zeros((6, pitch), dtype=dt)
        a.force = force.ctypes.data
        a.virial = vir.ctypes.data
        a.virial_pitch = pitch
        keep += [force, vir]
        return a, keep, force, vir

    def pair_forces(self, evaluator, table, pos, n_neigh, nlist, head, L, r_cut, ntypes=1,
                    r_on=0.0, mode="none", virial=True, tilt=(0, 0, 0), periodic=(1, 1, 1),
                    half=False, rint_image=False, nthreads=0, N=None):
        """``PotentialPair<E>::computeForces``. Returns (force (N,4), virial (6,N))."""
        a, keep, force, vir = self._args(pos, n_neigh, nlist, head, L, tilt, periodic, ntypes,
                                         r_cut, r_on, mode, virial, half, rint_image, nthreads, N)
        table = np.ascontiguousarray(table, dtype=np.uint8)
        rc = self.lib.oracle_pair_forces(EVALUATORS[evaluator], ctypes.byref(a),
                                         table.ctypes.data)
        assert rc == 0
        return force, vir

    def dpd_forces(self, table, pos, vel, tag, n_neigh, nlist, head, L, r_cut, seed, timestep,
                   dt, kT, ntypes=1, virial=True, tilt=(0, 0, 0), periodic=(1, 1, 1),
                   half=False, rint_image=False, nthreads=0, N=None):
        """``PotentialPairDPDThermo<GeneralWeight>::computeForces``."""
        a, keep, force, vir = self._args(pos, n_neigh, nlist, head, L, tilt, periodic, ntypes,
                                         r_cut, 0.0, "none", virial, half, rint_image, nthreads,
                                         N)
        vel = np.ascontiguousarray(vel, dtype=self.dtype)
        tag = np.ascontiguousarray(tag, dtype=np.uint32)
        a.vel = vel.ctypes.data
        a.tag = tag.ctypes.data
        a.seed = int(seed) & 0xFFFF
        a.timestep = int(timestep)
        a.deltaT = float(dt)
        a.T = float(kT)
        table = np.ascontiguousarray(table, dtype=np.uint8)
        rc = self.lib.oracle_dpd_forces(EVALUATORS["DPDGeneralWeight"], ctypes.byref(a),
                                        table.ctypes.data)
        assert rc == 0
        return force, vir

    def aniso_forces(self, table, pos, orientation, n_neigh, nlist, head, L, r_cut, ntypes=1,
                     mode="none", virial=True, tilt=(0, 0, 0), periodic=(1, 1, 1), half=False,
                     rint_image=False, nthreads=0, N=None):
        """``AnisoPotentialPair<TwoPatchMorse>::computeForces``. Returns (force, torque, virial)."""
        a, keep, force, vir = self._args(pos, n_neigh, nlist, head, L, tilt, periodic, ntypes,
                                         r_cut, 0.0, mode, virial, half, rint_image, nthreads, N)
        q = np.ascontiguousarray(orientation, dtype=self.dtype)
        torque = np.zeros((force.shape[0], 4), dtype=self.dtype)
        a.orientation = q.ctypes.data
        a.torque = torque.ctypes.data
        table = np.ascontiguousarray(table, dtype=np.uint8)
        rc = self.lib.oracle_aniso_forces(EVALUATORS["TwoPatchMorse"], ctypes.byref(a),
                                          table.ctypes.data)
        assert rc == 0
        return force, torque, vir


def particle_types(pos):
    """Type ids bit-cast in pos[:, 3] (HOOMD __scalar_as_int convention, Appendix A.1)."""
    pos = np.ascontiguousarray(pos)
    if pos.dtype == np.float32:
        return pos.view(np.uint32)[:, 3].copy()
    return (pos.view(np.uint64)[:, 3] & 0xFFFFFFFF).astype(np.uint32)


def make_pos(xyz, types, dtype):
    """(N,4) Scalar4 array with the type id bit-cast into .w."""
    xyz = np.asarray(xyz)
    N = xyz.shape[0]
    pos = np.zeros((N, 4), dtype=dtype)
    pos[:, :3] = xyz
    t = np.broadcast_to(np.asarray(types, dtype=np.uint32), (N,))
    if np.dtype(dtype) == np.float32:
        pos.view(np.uint32)[:, 3] = t
    else:
        pos.view(np.uint64)[:, 3] = t.astype(np.uint64)
    return pos
