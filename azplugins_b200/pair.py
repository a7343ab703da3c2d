"""Pair potentials -- the ``hoomd.azplugins.pair`` API surface on the B200 kernels.

Same class names, constructor arguments, ``params`` keys, ``r_cut`` / ``r_on`` / ``mode`` / ``kT``
semantics and C++ class-name mapping as reference ``src/pair.py`` (Colloid :14-118,
DPDGeneralWeight :121-239, ExpandedYukawa :242-297, Hertz :300-351, PerturbedLennardJones
:354-426, TwoPatchMorse :429-525), and the ``Force`` read-outs HOOMD exposes (``forces``,
``energies``, ``energy``, ``torques``, ``virials``). HOOMD itself is not required: a potential
attaches to an ``azplugins_b200.State`` and computes through the C ABI in
``include/azp_b200.h``. INTEGRATION.md shows how the same kernels slot under a real HOOMD build.

There is no CPU implementation behind these classes: ``hoomd.device.CPU``-style execution is
out of scope and attaching to a non-CUDA state raises.
"""

import numpy as np
import torch

from . import _lib, kernels
from .nlist import NeighborList


class _TypePairDict:
    """``TypeParameterDict(..., len_keys=2)`` work-alike: values keyed by unordered type pairs."""

    def __init__(self, schema=None, default=None, owner=None):
        self._schema = schema  # ordered {key: python type} or None for scalar entries
        self._default = default
        self._data = {}
        self._owner = owner  # the potential whose device tables go stale when a value changes

    def _touch(self):
        if self._owner is not None:
            self._owner._tables_version += 1

    @staticmethod
    def _key(key):
        if not (isinstance(key, tuple) and len(key) == 2):
            raise KeyError("type-pair keys are 2-tuples of type names, got %r" % (key,))
        a, b = key
        return (a, b) if a <= b else (b, a)

    def _validate(self, value):
        if self._schema is None:
            return None if value is None else float(value)
        if not isinstance(value, dict):
            raise TypeError("parameters must be given as a dict")
        missing = [k for k in self._schema if k not in value]
        extra = [k for k in value if k not in self._schema]
        if missing or extra:
            raise KeyError("parameter dict mismatch: missing %s, unexpected %s" % (missing, extra))
        out = {}
        for k, typ in self._schema.items():
            out[k] = bool(value[k]) if typ is bool else float(value[k])
        return out

    def __setitem__(self, key, value):
        if isinstance(key, tuple) and len(key) == 2 and any(isinstance(k, (list, tuple)) for k in key):
            # HOOMD allows (["A","B"], ["A","B"]) style multi-keys
            la = key[0] if isinstance(key[0], (list, tuple)) else [key[0]]
            lb = key[1] if isinstance(key[1], (list, tuple)) else [key[1]]
            for a in la:
                for b in lb:
                    self[(a, b)] = value
            return
        self._data[self._key(key)] = self._validate(value)
        self._touch()

    def __getitem__(self, key):
        k = self._key(key)
        if k in self._data:
            v = self._data[k]
            return dict(v) if isinstance(v, dict) else v
        if self._default is not None:
            return self._default
        raise KeyError("no value set for type pair %r" % (key,))

    def __contains__(self, key):
        return self._key(key) in self._data

    def keys(self):
        return self._data.keys()

    @property
    def default(self):
        return self._default

    @default.setter
    def default(self, v):
        self._default = self._validate(v)
        self._touch()


class Pair:
    """Base of the isotropic pair potentials (``hoomd.md.pair.Pair`` work-alike, Appendix A.9)."""

    _cpp_class_name = None
    _evaluator = None
    _family = _lib.FAMILY_PAIR
    _accepted_modes = ("none", "shift", "xplor")
    _param_schema = {}
    is_anisotropic = False

    def __init__(self, nlist, default_r_cut=None, default_r_on=0.0, mode="none"):
        if not isinstance(nlist, NeighborList):
            raise TypeError("nlist must be an azplugins_b200.nlist.NeighborList")
        self.nlist = nlist
        self._tables_version = 0   # bumped by every params / r_cut / r_on assignment
        self._uploaded_version = -1
        self.params = _TypePairDict(schema=dict(self._param_schema), owner=self)
        self.r_cut = _TypePairDict(default=None if default_r_cut is None else float(default_r_cut),
                                   owner=self)
        self.r_on = _TypePairDict(default=float(default_r_on), owner=self)
        self.mode = mode
        self._state = None
        self._launch_shape = (0, 0)  # (block_size, threads_per_particle); 0 = library default
        nlist._add_consumer(self)

    # ---- configuration -------------------------------------------------------------------
    @property
    def mode(self):
        return self._mode

    @mode.setter
    def mode(self, value):
        if value not in self._accepted_modes:
            raise ValueError("mode must be one of %s" % (self._accepted_modes,))
        self._mode = value

    @property
    def cpp_class_name(self):
        """Name of the C++ class HOOMD would instantiate on a GPU device (``+"GPU"``)."""
        return self._cpp_class_name + "GPU"

    @property
    def kernel_parameters(self):
        """(block_size, threads_per_particle) -- HOOMD's Autotuner<2> dimensions."""
        return self._launch_shape

    @kernel_parameters.setter
    def kernel_parameters(self, value):
        self._launch_shape = (int(value[0]), int(value[1]))

    def _r_cut_matrix(self, state):
        nt = state.ntypes
        rc = np.zeros((nt, nt))
        for i, a in enumerate(state.types):
            for j, b in enumerate(state.types):
                v = self.r_cut[(a, b)] if ((a, b) in self.r_cut or self.r_cut.default is not None) else None
                if v is None:
                    raise ValueError("r_cut not set for type pair (%s, %s)" % (a, b))
                rc[i, j] = float(v)
        return rc

    def _r_on_matrix(self, state):
        nt = state.ntypes
        ro = np.zeros((nt, nt))
        for i, a in enumerate(state.types):
            for j, b in enumerate(state.types):
                ro[i, j] = float(self.r_on[(a, b)])
        return ro

    def _fields(self, p):
        return [float(p[k]) for k in self._param_schema]

    # ---- attach: pack param_type tables and upload them (call stack 3.1 step 4) ----------
    def attach(self, state):
        if state.device.type != "cuda":
            raise _lib.AzpError("%s runs on CUDA devices only (no CPU fallback)" % type(self).__name__)
        self._state = state
        self._upload_tables()
        dev = state.device
        n = state.N
        self._force = torch.zeros((n, 4), dtype=state.torch_dtype, device=dev)
        self._virial = torch.zeros((6, n), dtype=state.torch_dtype, device=dev)
        self._torque = torch.zeros((n, 4), dtype=state.torch_dtype, device=dev)
        self._computed_at = None
        return self

    def _upload_tables(self):
        """Pack and upload param_type / r_cut^2 / r_on^2 tables (HOOMD re-syncs them whenever a
        parameter is set; here: at attach and before the next compute after any assignment)."""
        state = self._state
        bits = 8 * state.dtype.itemsize
        nt = state.ntypes
        psz = kernels.param_size(self._evaluator, bits)
        table = np.zeros((nt * nt, psz), dtype=np.uint8)
        for i, a in enumerate(state.types):
            for j, b in enumerate(state.types):
                if (a, b) not in self.params:
                    raise ValueError("params not set for type pair (%s, %s)" % (a, b))
                table[j * nt + i] = kernels.pack_params(self._evaluator, bits, self._fields(self.params[(a, b)]))
        rc = self._r_cut_matrix(state).astype(state.dtype)
        ro = self._r_on_matrix(state).astype(state.dtype)
        dev = state.device
        self._bits = bits
        self._d_params = torch.from_numpy(table.reshape(-1).copy()).to(dev)
        self._d_rcutsq = torch.from_numpy((rc * rc).T.reshape(-1).copy()).to(dev)
        self._d_ronsq = torch.from_numpy((ro * ro).T.reshape(-1).copy()).to(dev)
        self._uploaded_version = self._tables_version

    _attach_hook = attach

    def get_params_from_device(self, type_a, type_b):
        """asDict()/toPython() of the uploaded ``param_type`` (what HOOMD reads back after attach)."""
        st = self._state
        nt = st.ntypes
        psz = kernels.param_size(self._evaluator, self._bits)
        idx = st.type_index(type_b) * nt + st.type_index(type_a)
        raw = self._d_params.view(nt * nt, psz)[idx].cpu().numpy()
        f = kernels.unpack_params(self._evaluator, self._bits, raw)
        out = {}
        for k, v in zip(self._param_schema, f):
            out[k] = bool(v) if self._param_schema[k] is bool else float(v)
        return out

    # ---- compute -------------------------------------------------------------------------
    def _extra_args(self, timestep):
        return {}

    def _args(self, timestep=None, compute_virial=True, row_ids=None, rows=None):
        st = self._state
        if st is None:
            raise RuntimeError("potential is not attached to a State")
        if self._uploaded_version != self._tables_version:
            self._upload_tables()
        self.nlist.compute(st)
        if self.nlist.storage_mode != "full":
            raise RuntimeError("GPU pair potentials need a full neighbour list")
        ts = st.timestep if timestep is None else int(timestep)
        lo, hi = (0, st.N) if rows is None else (int(rows[0]), int(rows[1]))
        if not (0 <= lo <= hi <= st.N):
            raise ValueError("rows must satisfy 0 <= lo <= hi <= N")
        extra = self._extra_args(ts)
        if "torque" in extra:
            extra["torque"] = extra["torque"][lo:hi]
        # a contiguous row range is the same launch on offset views of the per-row arrays
        # (n_neigh, head_list, outputs); per-particle inputs stay whole and are indexed with
        # row_offset, the virial keeps its full pitch
        return kernels.fill_args(
            box=st.box, pos=st.pos, n_neigh=self.nlist.n_neigh[lo:hi], nlist=self.nlist.nlist,
            head_list=self.nlist.head_list[lo:hi], rcutsq=self._d_rcutsq, ronsq=self._d_ronsq,
            ntypes=st.ntypes, force=self._force[lo:hi],
            virial=self._virial[:, lo:hi] if compute_virial else None,
            virial_pitch=self._virial.shape[1],
            n_rows=hi - lo, row_offset=lo, shift_mode=_lib.SHIFT_MODES[self._mode],
            compute_virial=compute_virial,
            block_size=self._launch_shape[0], threads_per_particle=self._launch_shape[1],
            timestep=ts, size_neigh_list=self.nlist.size, row_ids=row_ids,
            n_max=self.nlist.n_max,
            **extra)

    def _args_all_rows(self, timestep, compute_virial):
        """The argument struct of a launch over all rows, rebuilt only when something it points
        at has changed. Filling ``azp_pair_args`` field by field from Python (tensor views, box
        conversion, ~30 ctypes stores) costs more host time than the C1 kernel runs (N = 32,000:
        0.015 ms); a time step that changes nothing but ``timestep`` reuses the struct. Potentials
        with per-step extra arguments (DPD: kT(t), velocities; aniso: orientations) always take
        the full path."""
        if type(self)._extra_args is not Pair._extra_args:
            return self._args(timestep, compute_virial)
        st = self._state
        if st is None:
            raise RuntimeError("potential is not attached to a State")
        if self._uploaded_version != self._tables_version:
            self._upload_tables()
        nl = self.nlist
        nl.compute(st)
        b = st.box
        key = (st.pos.data_ptr(), nl.n_neigh.data_ptr(), nl.nlist.data_ptr(), nl.head_list.data_ptr(),
               nl.size, nl.n_max, bool(compute_virial), self._launch_shape, self._mode,
               self._uploaded_version, st.N, self._force.data_ptr(), b.Lx, b.Ly, b.Lz, b.xy, b.xz,
               b.yz, b.periodic)
        cached = getattr(self, "_args_cache", None)
        if cached is None or cached[0] != key:
            cached = (key, self._args(timestep, compute_virial))
            self._args_cache = cached
        args = cached[1]
        args.timestep = st.timestep if timestep is None else int(timestep)
        return args

    def compute(self, timestep=None, compute_virial=True, row_ids=None, rows=None):
        """``ForceCompute::compute(timestep)``: enqueue the kernel on the current stream.
        ``row_ids`` (int32 device tensor) restricts the evaluation to those rows (scheduler use:
        interior rows while the halo exchange is in flight, boundary rows after); ``rows`` =
        ``(lo, hi)`` to a contiguous range (used by :meth:`compute_to_host`)."""
        if row_ids is not None and rows is not None:
            raise ValueError("give row_ids or rows, not both")
        if row_ids is None and rows is None:
            args = self._args_all_rows(timestep, compute_virial)
        else:
            args = self._args(timestep, compute_virial, row_ids, rows)
        if args.N == 0:
            return self
        dev = self._state.device
        if torch.cuda.current_device() == dev.index:
            kernels.launch(self._family, self._evaluator, self._bits, args, self._d_params.data_ptr())
        else:
            with torch.cuda.device(dev):
                kernels.launch(self._family, self._evaluator, self._bits, args, self._d_params.data_ptr())
        if row_ids is None and rows is None:
            self._computed_at = (self._state.timestep if timestep is None else int(timestep),
                                 self._tables_version)
        return self

    def compute_to_host(self, host_force, host_virial=None, host_torque=None, timestep=None,
                        chunks=4):
        """Evaluate and deliver the per-particle results into pinned host tensors, overlapping the
        device-to-host copies with the computation: the rows are evaluated in ``chunks``
        contiguous ranges and every finished range is copied on a second stream while the next
        one runs (the copy of 40 MB of forces + virials takes 2.5x the C2 kernel, so the step is
        bounded by PCIe, not by kernel + copy). Synchronises before returning."""
        st = self._state
        if st is None:
            raise RuntimeError("potential is not attached to a State")
        compute_virial = host_virial is not None
        cur = torch.cuda.current_stream(st.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=st.device)
            self._chunk_events = []
        while len(self._chunk_events) < chunks:
            self._chunk_events.append(torch.cuda.Event())
        n = st.N
        bounds = [(n * c) // chunks for c in range(chunks + 1)]
        for c in range(chunks):
            lo, hi = bounds[c], bounds[c + 1]
            if hi == lo:
                continue
            self.compute(timestep=timestep, compute_virial=compute_virial, rows=(lo, hi))
            ev = self._chunk_events[c]
            ev.record(cur)
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(ev)
                host_force[lo:hi].copy_(self._force[lo:hi], non_blocking=True)
                if compute_virial:
                    for k in range(6):  # six contiguous segments of the (6, N) array
                        host_virial[k, lo:hi].copy_(self._virial[k, lo:hi], non_blocking=True)
                if host_torque is not None:
                    host_torque[lo:hi].copy_(self._torque[lo:hi], non_blocking=True)
        self._copy_stream.synchronize()
        return self

    def tune_kernel_parameters(self, timestep=None, compute_virial=True):
        """Scan (block_size, threads_per_particle) like HOOMD's autotuner and pin the fastest."""
        args = self._args(timestep, compute_virial)
        with torch.cuda.device(self._state.device):
            b, t, ms = kernels.autotune(self._family, self._evaluator, self._bits, args,
                                        self._d_params.data_ptr())
        self._launch_shape = (b, t)
        return b, t, ms

    # ---- read-outs (hoomd.md.force.Force) ------------------------------------------------------
    def _need(self):
        # ForceCompute::compute(timestep) semantics: evaluate once per time step (and again after
        # a parameter change); an explicit compute() always re-evaluates
        if getattr(self, "_computed_at", None) != (self._state.timestep, self._tables_version):
            self.compute()

    @property
    def forces(self):
        self._need()
        return self._force[:, :3].cpu().numpy()

    @property
    def energies(self):
        self._need()
        return self._force[:, 3].cpu().numpy()

    @property
    def energy(self):
        self._need()
        return float(self._force[:, 3].sum(dtype=torch.float64).item())

    @property
    def torques(self):
        self._need()
        return self._torque[:, :3].cpu().numpy()

    @property
    def virials(self):
        self._need()
        return self._virial.cpu().numpy().T.copy()


class Colloid(Pair):
    """Colloid (integrated Lennard-Jones) potential; params ``A, a_1, a_2, sigma``."""

    _cpp_class_name = "PotentialPairColloid"
    _evaluator = _lib.EV_COLLOID
    _param_schema = {"A": float, "a_1": float, "a_2": float, "sigma": float}


class ExpandedYukawa(Pair):
    """Expanded Yukawa potential; params ``epsilon, kappa, delta``."""

    _cpp_class_name = "PotentialPairExpandedYukawa"
    _evaluator = _lib.EV_EXPANDED_YUKAWA
    _param_schema = {"epsilon": float, "kappa": float, "delta": float}


class Hertz(Pair):
    """Hertz potential; params ``epsilon``."""

    _cpp_class_name = "PotentialPairHertz"
    _evaluator = _lib.EV_HERTZ
    _param_schema = {"epsilon": float}


class PerturbedLennardJones(Pair):
    """Perturbed Lennard-Jones; params ``epsilon, sigma, attraction_scale_factor``."""

    _cpp_class_name = "PotentialPairPerturbedLennardJones"
    _evaluator = _lib.EV_PERTURBED_LENNARD_JONES
    _param_schema = {"epsilon": float, "sigma": float, "attraction_scale_factor": float}


class DPDGeneralWeight(Pair):
    """DPD with generalised weight function and thermostat; params ``A, gamma, s``; ``kT``."""

    _cpp_class_name = "PotentialPairDPDThermoGeneralWeight"
    _evaluator = _lib.EV_DPD_GENERAL_WEIGHT
    _family = _lib.FAMILY_DPD
    _accepted_modes = ("none",)
    _param_schema = {"A": float, "gamma": float, "s": float}

    def __init__(self, nlist, kT, default_r_cut=None):
        super().__init__(nlist=nlist, default_r_cut=default_r_cut, default_r_on=0, mode="none")
        self.kT = kT

    def _kT(self, timestep):
        return float(self.kT(timestep)) if callable(self.kT) else float(self.kT)

    def _extra_args(self, timestep):
        st = self._state
        return dict(vel=st.vel, tag=st.tag, seed=st.seed, dt=st.dt, kT=self._kT(timestep))


class DPDGeneralWeightConservative(DPDGeneralWeight):
    """``PotentialPairConservativeGeneralWeight``: the conservative part only
    (reference src/export_PotentialPairDPDThermo.cc.inc:33-35)."""

    _cpp_class_name = "PotentialPairConservativeGeneralWeight"
    _family = _lib.FAMILY_PAIR

    def __init__(self, nlist, default_r_cut=None):
        super().__init__(nlist=nlist, kT=0.0, default_r_cut=default_r_cut)

    def _extra_args(self, timestep):
        return {}


class AnisotropicPair(Pair):
    """``hoomd.md.pair.aniso.AnisotropicPair`` work-alike: modes none/shift, no r_on."""

    _family = _lib.FAMILY_ANISO
    _accepted_modes = ("none", "shift")
    is_anisotropic = True

    def __init__(self, nlist, default_r_cut=None, mode="none"):
        super().__init__(nlist, default_r_cut, 0.0, mode)

    def _extra_args(self, timestep):
        return dict(orientation=self._state.orientation, torque=self._torque)


class TwoPatchMorse(AnisotropicPair):
    """Two-patch Morse; params ``M_d, M_r, r_eq, omega, alpha, repulsion``."""

    _cpp_class_name = "AnisoPotentialPairTwoPatchMorse"
    _evaluator = _lib.EV_TWO_PATCH_MORSE
    _param_schema = {"M_d": float, "M_r": float, "r_eq": float, "omega": float,
                     "alpha": float, "repulsion": bool}


class FusedPair:
    """Two attached isotropic potentials on ONE neighbour list evaluated in one sweep of the list
    (C ABI ``azp_pair_forces_fused_*``). The reference's documented case is
    ``Colloid`` + ``Hertz`` (reference src/pair.py:66-76), which HOOMD runs as two force computes
    that each stream the list. Each potential keeps its own parameters, cutoffs, mode (none /
    shift) and read-outs; after :meth:`compute` both hold exactly what their own ``compute()``
    would have produced."""

    SUPPORTED = {(_lib.EV_COLLOID, _lib.EV_HERTZ)}

    def __init__(self, a, b):
        if (a._evaluator, b._evaluator) not in self.SUPPORTED:
            if (b._evaluator, a._evaluator) in self.SUPPORTED:
                a, b = b, a
            else:
                raise ValueError("no fused kernel for (%s, %s)" % (type(a).__name__, type(b).__name__))
        if a.nlist is not b.nlist:
            raise ValueError("fused potentials must share one neighbour list")
        if "xplor" in (a.mode, b.mode):
            raise ValueError("xplor potentials are evaluated separately")
        self.a, self.b = a, b

    @staticmethod
    def can_fuse(a, b):
        pair_ok = (a._evaluator, b._evaluator) in FusedPair.SUPPORTED or \
            (b._evaluator, a._evaluator) in FusedPair.SUPPORTED
        return (pair_ok and a.nlist is b.nlist and "xplor" not in (a.mode, b.mode)
                and a._family == _lib.FAMILY_PAIR and b._family == _lib.FAMILY_PAIR)

    @property
    def kernel_parameters(self):
        return self.a.kernel_parameters

    @kernel_parameters.setter
    def kernel_parameters(self, value):
        self.a.kernel_parameters = value

    def compute(self, timestep=None, compute_virial=True, row_ids=None, rows=None):
        a, b = self.a, self.b
        if a._state is None or a._state is not b._state:
            raise RuntimeError("fused potentials must be attached to the same State")
        args_a = a._args(timestep, compute_virial, row_ids, rows)
        args_b = b._args(timestep, compute_virial, row_ids, rows)
        if args_a.N == 0:
            return self
        with torch.cuda.device(a._state.device):
            kernels.launch_fused(a._evaluator, args_a, a._d_params.data_ptr(), b._evaluator, args_b,
                                 b._d_params.data_ptr(), a._bits)
        if row_ids is None and rows is None:
            ts = a._state.timestep if timestep is None else int(timestep)
            a._computed_at = (ts, a._tables_version)
            b._computed_at = (ts, b._tables_version)
        return self

    def tune_kernel_parameters(self, timestep=None, compute_virial=True, reps=3):
        """Scan (block_size, threads_per_particle) for the fused launch and pin the fastest."""
        best = None
        for block in (64, 128, 256):
            for tpp in (1, 2, 4, 8):
                self.kernel_parameters = (block, tpp)
                self.compute(timestep, compute_virial)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    self.compute(timestep, compute_virial)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                if best is None or ms < best[2]:
                    best = (block, tpp, ms)
        self.kernel_parameters = best[:2]
        return best
