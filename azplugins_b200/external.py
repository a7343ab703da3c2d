"""``hoomd.azplugins.external`` on the B200 path: the external harmonic barriers.

Mirrors reference ``src/external.py:14-155`` (class names, ``location`` variant, per-type
``params`` with keys ``k`` and ``offset``, C++ class name + ``"GPU"`` on a GPU device) on top of
the C ABI ``azp_harmonic_barrier_f32/_f64`` (``include/azp_b200.h``), which replaces
``gpu::compute_harmonic_barrier<BarrierEvaluatorT>`` (reference src/HarmonicBarrierGPU.cuh:100-132).
Like the reference, the virial contribution is not computed (the array is zeroed) and an invalid
barrier position raises ``RuntimeError("Barrier position is invalid")``
(src/HarmonicBarrier.h:126-130). CUDA only: there is no CPU fallback.
"""

import ctypes

import numpy as np
import torch

from . import _lib


class _TypeParams(dict):
    """``TypeParameter`` stand-in keyed by particle type name; values are dicts ``{k, offset}``."""

    def __setitem__(self, key, value):
        value = dict(value)
        if set(value) != {"k", "offset"}:
            raise ValueError("params must have exactly the keys 'k' and 'offset'")
        super().__setitem__(key, dict(k=float(value["k"]), offset=float(value["offset"])))


class HarmonicBarrier:
    """Repulsive barrier implemented as a harmonic potential (reference src/external.py:14-84).
    Use a derived type. ``location``: float or callable ``location(timestep) -> float``
    (``hoomd.variant.variant_like``)."""

    _geometry = None

    def __init__(self, location):
        if self._geometry is None:
            raise TypeError("HarmonicBarrier should not be instantiated directly; use a derived type")
        self.location = location
        self.params = _TypeParams()
        self._state = None
        self.block_size = 0

    @property
    def location(self):
        return self._location

    @location.setter
    def location(self, value):
        if callable(value):
            self._location = value
        else:
            v = float(value)
            self._location = _Constant(v)

    @property
    def cpp_class_name(self):
        """Name of the C++ class HOOMD would instantiate on a GPU device (src/external.py:75-79)."""
        return type(self).__name__ + "GPU"

    def attach(self, state):
        if state.device.type != "cuda":
            raise _lib.AzpError("%s runs on CUDA devices only (no CPU fallback)" % type(self).__name__)
        self._state = state
        n = state.N
        self._force = torch.zeros((n, 4), dtype=state.torch_dtype, device=state.device)
        self._virial = torch.zeros((6, n), dtype=state.torch_dtype, device=state.device)
        self._d_params = None
        self._params_key = None
        self._computed = False
        return self

    def _upload_params(self):
        st = self._state
        table = np.zeros((st.ntypes, 2), dtype=st.dtype)
        for i, name in enumerate(st.types):
            if name not in self.params:
                raise ValueError("params not set for particle type %s" % name)
            table[i] = (self.params[name]["k"], self.params[name]["offset"])
        key = table.tobytes()
        if key != self._params_key:
            self._d_params = torch.from_numpy(table).to(st.device)
            self._params_key = key

    def compute(self, timestep=None):
        """``ForceCompute::compute(timestep)``: evaluate the barrier at ``location(timestep)``."""
        st = self._state
        if st is None:
            raise RuntimeError("barrier is not attached to a State")
        ts = st.timestep if timestep is None else int(timestep)
        loc = float(self._location(ts))
        bits = 8 * st.dtype.itemsize
        box = st.box.to_c()
        if not _lib.lib.azp_harmonic_barrier_valid(self._geometry, bits, loc, ctypes.byref(box)):
            raise RuntimeError("Barrier position is invalid")
        self._upload_params()
        a = _lib.AzpBarrierArgs()
        a.d_force = self._force.data_ptr()
        a.d_virial = self._virial.data_ptr()
        a.virial_pitch = self._virial.shape[1]
        a.d_pos = st.pos.data_ptr()
        a.d_params = self._d_params.data_ptr()
        a.box = box
        a.location = loc
        a.N = st.N
        a.ntypes = st.ntypes
        a.geometry = self._geometry
        a.block_size = int(self.block_size)
        with torch.cuda.device(st.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            rc = getattr(_lib.lib, "azp_harmonic_barrier_f%d" % bits)(ctypes.byref(a), stream)
        _lib.check(rc, "harmonic barrier launch")
        self._computed = True
        return self

    # ---- read-outs (hoomd.md.force.Force) -------------------------------------------------
    def _need(self):
        if not getattr(self, "_computed", False):
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
    def virials(self):
        self._need()
        return self._virial.cpu().numpy().T.copy()


class _Constant:
    """``hoomd.variant.Constant`` stand-in."""

    def __init__(self, value):
        self.value = value

    def __call__(self, timestep):
        return self.value


class PlanarHarmonicBarrier(HarmonicBarrier):
    """Planar harmonic barrier normal to *y* at ``y = location`` (reference src/external.py:87-120,
    src/PlanarBarrierEvaluator.h): ``U = k/2 (y - H)^2`` for ``y > H = location + offset``."""

    _geometry = _lib.BARRIER_PLANAR


class SphericalHarmonicBarrier(HarmonicBarrier):
    """Spherical harmonic barrier of radius ``location`` about the origin (reference
    src/external.py:123-155, src/SphericalBarrierEvaluator.h): ``U = k/2 (r - R)^2`` for
    ``r > R = location + offset``."""

    _geometry = _lib.BARRIER_SPHERICAL
