"""``hoomd.azplugins.wall`` on the B200 path: the Colloid and LJ 9-3 wall potentials.

Mirrors reference ``src/wall.py:14-146`` (class names, ``walls`` list, per-type ``params`` with
keys ``A``, ``a``, ``sigma``, ``r_cut``, ``r_extrap``; C++ class names ``WallsPotentialColloid`` /
``WallsPotentialLJ93`` + ``"GPU"``) and the ``hoomd.wall`` geometries it is given (``Sphere``,
``Cylinder``, ``Plane``), on top of the C ABI ``azp_wall_forces_f32/_f64``
(``include/azp_b200.h``). The evaluator arithmetic is the reference's
(``src/WallEvaluatorColloid.h``, ``src/WallEvaluatorLJ93.h``); the wall loop restates HOOMD's
``EvaluatorWalls`` (not in the reference tree), including its extrapolated mode (``r_extrap > 0``:
closer to a wall than ``r_extrap`` the potential continues linearly). CUDA only: there is no CPU
fallback.
"""

import ctypes

import numpy as np
import torch

from . import _lib

MAX_SPHERES, MAX_CYLINDERS, MAX_PLANES = 20, 20, 60


class Sphere:
    """``hoomd.wall.Sphere(radius, origin=(0,0,0), inside=True, open=True)``."""

    def __init__(self, radius, origin=(0.0, 0.0, 0.0), inside=True, open=True):
        self.radius, self.origin = float(radius), tuple(float(v) for v in origin)
        self.inside, self.open = bool(inside), bool(open)


class Cylinder:
    """``hoomd.wall.Cylinder(radius, axis, origin=(0,0,0), inside=True, open=True)``."""

    def __init__(self, radius, axis, origin=(0.0, 0.0, 0.0), inside=True, open=True):
        self.radius, self.origin = float(radius), tuple(float(v) for v in origin)
        ax = np.asarray(axis, dtype=np.float64)
        n = np.linalg.norm(ax)
        if n == 0:
            raise ValueError("cylinder axis must not be zero")
        self.axis = tuple(ax / n)
        self.inside, self.open = bool(inside), bool(open)


class Plane:
    """``hoomd.wall.Plane(origin, normal, open=True)``; the normal points into the active space."""

    def __init__(self, origin, normal, open=True):
        self.origin = tuple(float(v) for v in origin)
        nv = np.asarray(normal, dtype=np.float64)
        n = np.linalg.norm(nv)
        if n == 0:
            raise ValueError("plane normal must not be zero")
        self.normal = tuple(nv / n)
        self.open = bool(open)


def _walls_struct(scalar):
    class _Sphere(ctypes.Structure):
        _fields_ = [("r", scalar), ("o", scalar * 3), ("inside", ctypes.c_int32), ("open", ctypes.c_int32)]

    class _Cylinder(ctypes.Structure):
        _fields_ = [("r", scalar), ("o", scalar * 3), ("a", scalar * 3), ("inside", ctypes.c_int32),
                    ("open", ctypes.c_int32)]

    class _Plane(ctypes.Structure):
        _fields_ = [("o", scalar * 3), ("n", scalar * 3), ("open", ctypes.c_int32), ("_pad", ctypes.c_int32)]

    class _Walls(ctypes.Structure):
        _fields_ = [("n_spheres", ctypes.c_uint32), ("n_cylinders", ctypes.c_uint32),
                    ("n_planes", ctypes.c_uint32), ("_pad", ctypes.c_uint32),
                    ("spheres", _Sphere * MAX_SPHERES), ("cylinders", _Cylinder * MAX_CYLINDERS),
                    ("planes", _Plane * MAX_PLANES)]

    return _Walls


def pack_walls(walls, dtype):
    """Bytes of the library's wall list (``d_walls`` of ``azp_wall_args``) for ``walls``."""
    bits = 8 * np.dtype(dtype).itemsize
    W = _walls_struct(ctypes.c_float if bits == 32 else ctypes.c_double)
    assert ctypes.sizeof(W) == _lib.lib.azp_walls_size(bits), "wall list layout mismatch"
    w = W()
    for g in walls:
        if isinstance(g, Sphere):
            if w.n_spheres >= MAX_SPHERES:
                raise ValueError("at most %d sphere walls" % MAX_SPHERES)
            s = w.spheres[w.n_spheres]
            s.r, s.inside, s.open = g.radius, int(g.inside), int(g.open)
            s.o[:] = g.origin
            w.n_spheres += 1
        elif isinstance(g, Cylinder):
            if w.n_cylinders >= MAX_CYLINDERS:
                raise ValueError("at most %d cylinder walls" % MAX_CYLINDERS)
            c = w.cylinders[w.n_cylinders]
            c.r, c.inside, c.open = g.radius, int(g.inside), int(g.open)
            c.o[:] = g.origin
            c.a[:] = g.axis
            w.n_cylinders += 1
        elif isinstance(g, Plane):
            if w.n_planes >= MAX_PLANES:
                raise ValueError("at most %d plane walls" % MAX_PLANES)
            p = w.planes[w.n_planes]
            p.o[:] = g.origin
            p.n[:] = g.normal
            p.open = int(g.open)
            w.n_planes += 1
        else:
            raise TypeError("walls must be wall.Sphere, wall.Cylinder or wall.Plane")
    return bytes(w)


def walls_as_arrays(walls):
    """(spheres, cylinders, planes) as plain rows -- the form the test oracle takes."""
    sph = [[g.radius, *g.origin, g.inside, g.open] for g in walls if isinstance(g, Sphere)]
    cyl = [[g.radius, *g.origin, *g.axis, g.inside, g.open] for g in walls if isinstance(g, Cylinder)]
    pla = [[*g.origin, *g.normal, g.open] for g in walls if isinstance(g, Plane)]
    return sph, cyl, pla


class _TypeParams(dict):
    def __init__(self, keys, defaults):
        super().__init__()
        self._keys, self._defaults = keys, defaults

    def __setitem__(self, key, value):
        v = dict(self._defaults)
        v.update(value)
        if set(v) != set(self._keys):
            raise ValueError("params must have the keys %s" % sorted(self._keys))
        super().__setitem__(key, {k: float(v[k]) for k in self._keys})


class WallPotential:
    """``hoomd.md.external.wall.WallPotential`` work-alike."""

    _evaluator = None
    _cpp_class_name = None
    _param_keys = ()

    def __init__(self, walls):
        if self._evaluator is None:
            raise TypeError("use wall.Colloid or wall.LJ93")
        self.walls = list(walls)
        self.params = _TypeParams(self._param_keys, {"r_extrap": 0.0})
        self._state = None
        self.block_size = 0

    @property
    def cpp_class_name(self):
        return self._cpp_class_name + "GPU"

    def _row(self, p, dtype):
        raise NotImplementedError

    def param_table(self, types, dtype):
        """Per-type rows of the library's ``d_params`` (the roundings of the reference
        constructors are reproduced in ``dtype``)."""
        rows = []
        for name in types:
            if name not in self.params:
                raise ValueError("params not set for particle type %s" % name)
            p = self.params[name]
            if p["r_extrap"] < 0.0:
                raise ValueError("r_extrap must be >= 0")
            S = np.dtype(dtype).type
            rc = S(p["r_cut"])
            rows.append(list(self._row(p, S)) + [rc * rc, S(p["r_extrap"])])
        return np.asarray(rows, dtype=dtype)

    def attach(self, state):
        if state.device.type != "cuda":
            raise _lib.AzpError("%s runs on CUDA devices only (no CPU fallback)" % type(self).__name__)
        self._state = state
        n = state.N
        self._force = torch.zeros((n, 4), dtype=state.torch_dtype, device=state.device)
        self._virial = torch.zeros((6, n), dtype=state.torch_dtype, device=state.device)
        self._key = None
        self._computed = False
        return self

    def _upload(self):
        st = self._state
        table = self.param_table(st.types, st.dtype)
        bits = 8 * st.dtype.itemsize
        assert table.shape[1] * st.dtype.itemsize == _lib.lib.azp_wall_param_size(self._evaluator, bits)
        blob = pack_walls(self.walls, st.dtype)
        key = (table.tobytes(), blob)
        if key != self._key:
            self._d_params = torch.from_numpy(table).to(st.device)
            self._d_walls = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(st.device)
            self._key = key

    def compute(self, timestep=None):
        """``ForceCompute::compute(timestep)``: enqueue the kernel on the current stream."""
        st = self._state
        if st is None:
            raise RuntimeError("wall potential is not attached to a State")
        self._upload()
        a = _lib.AzpWallArgs()
        a.d_force = self._force.data_ptr()
        a.d_virial = self._virial.data_ptr()
        a.virial_pitch = self._virial.shape[1]
        a.d_pos = st.pos.data_ptr()
        a.d_params = self._d_params.data_ptr()
        a.d_walls = self._d_walls.data_ptr()
        a.N = st.N
        a.ntypes = st.ntypes
        a.block_size = int(self.block_size)
        bits = 8 * st.dtype.itemsize
        with torch.cuda.device(st.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            rc = getattr(_lib.lib, "azp_wall_forces_f%d" % bits)(self._evaluator, ctypes.byref(a), stream)
        _lib.check(rc, "wall potential launch")
        self._computed = True
        return self

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


class Colloid(WallPotential):
    """Colloid wall potential (reference src/wall.py:14-81, src/WallEvaluatorColloid.h)."""

    _evaluator = _lib.WALL_COLLOID
    _cpp_class_name = "WallsPotentialColloid"
    _param_keys = ("A", "a", "sigma", "r_cut", "r_extrap")

    def _row(self, p, S):
        # src/WallEvaluatorColloid.h:36-45: c_1 = A sigma^6 / 7560, c_2 = A / 6
        A, sigma = S(p["A"]), S(p["sigma"])
        sigma_3 = sigma * sigma * sigma
        return A * sigma_3 * sigma_3 / S(7560), A / S(6), S(p["a"])


class LJ93(WallPotential):
    """Lennard-Jones 9-3 wall potential (reference src/wall.py:84-146, src/WallEvaluatorLJ93.h)."""

    _evaluator = _lib.WALL_LJ93
    _cpp_class_name = "WallsPotentialLJ93"
    _param_keys = ("A", "sigma", "r_cut", "r_extrap")

    def _row(self, p, S):
        # src/WallEvaluatorLJ93.h:37-43: sigma_3 = sigma^3
        sigma = S(p["sigma"])
        return sigma * sigma * sigma, S(p["A"])
