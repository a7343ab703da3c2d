"""Device-resident particle data in HOOMD's ``ParticleData`` layouts.

``pos``  Scalar4 (x, y, z, type id bit-cast)     ``vel``  Scalar4 (vx, vy, vz, mass)
``orientation`` Scalar4 quaternion (s, x, y, z)  ``tag``  uint32 global id (stored as int32 bits)

PyTorch is used only as the owner of device memory and streams.
"""

import numpy as np
import torch

from .box import Box


def _torch_dtype(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return torch.float32
    if dtype == np.float64:
        return torch.float64
    raise ValueError("Scalar must be float32 or float64")


def pack_pos(xyz, typeid, dtype):
    """numpy (N,4) Scalar4 with the type id bit-cast into .w (``__int_as_scalar``)."""
    xyz = np.asarray(xyz)
    n = xyz.shape[0]
    pos = np.zeros((n, 4), dtype=dtype)
    pos[:, :3] = xyz
    t = np.broadcast_to(np.asarray(typeid, dtype=np.uint32), (n,))
    if np.dtype(dtype) == np.float32:
        pos.view(np.uint32)[:, 3] = t
    else:
        pos.view(np.uint64)[:, 3] = t.astype(np.uint64)
    return pos


class State:
    """Particles + box + run metadata (seed, timestep, dt) the force computes read."""

    def __init__(self, box, types, position, typeid=None, velocity=None, mass=None,
                 orientation=None, tag=None, dtype=np.float32, device="cuda:0", seed=0,
                 timestep=0, dt=0.005, n_ghost=0):
        if not isinstance(box, Box):
            raise TypeError("box must be an azplugins_b200.Box")
        self.box = box
        self.types = list(types)
        self.dtype = np.dtype(dtype)
        self.device = torch.device(device)
        self.seed = int(seed) & 0xFFFF
        self.timestep = int(timestep)
        self.dt = float(dt)
        position = np.asarray(position, dtype=np.float64)
        n_total = position.shape[0]
        self.N = n_total - int(n_ghost)
        self.n_ghost = int(n_ghost)
        typeid = np.zeros(n_total, dtype=np.uint32) if typeid is None else np.asarray(typeid)
        if typeid.size and int(typeid.max()) >= len(self.types):
            raise ValueError("typeid out of range")
        self.pos = self._to_device(pack_pos(position, typeid, self.dtype))
        vel = np.zeros((n_total, 4), dtype=self.dtype)
        if velocity is not None:
            vel[:, :3] = np.asarray(velocity)
        vel[:, 3] = 1.0 if mass is None else np.asarray(mass)
        self.vel = self._to_device(vel)
        q = np.zeros((n_total, 4), dtype=self.dtype)
        q[:, 0] = 1.0
        if orientation is not None:
            q[:] = np.asarray(orientation)
        self.orientation = self._to_device(q)
        tag = np.arange(n_total, dtype=np.uint32) if tag is None else np.asarray(tag, dtype=np.uint32)
        self.tag = self._to_device(tag.view(np.int32))

    def _to_device(self, arr):
        t = torch.from_numpy(np.ascontiguousarray(arr))
        return t.to(self.device)

    @property
    def ntypes(self):
        return len(self.types)

    @property
    def torch_dtype(self):
        return _torch_dtype(self.dtype)

    def type_index(self, name):
        return self.types.index(name)

    def sfc_sort(self):
        """Reorder the particles along a Morton curve on the device (what HOOMD's SFCPackTuner
        does to ParticleData): ``pos``, ``vel``, ``orientation`` and ``tag`` are permuted together,
        so spatial neighbours become memory neighbours and the position gathers of the force
        kernels stay L1/L2-resident. Returns the permutation (new index -> old index). Neighbour
        lists and per-particle outputs refer to the old order: rebuild / recompute afterwards."""
        import ctypes

        from . import _lib

        n = self.pos.shape[0] - self.n_ghost
        order = torch.empty(n, dtype=torch.int32, device=self.device)
        box = self.box.to_c()
        with torch.cuda.device(self.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            fn = getattr(_lib.lib, "azp_sfc_order_f%d" % (8 * self.dtype.itemsize))
            _lib.check(fn(self.pos.data_ptr(), ctypes.byref(box), n, order.data_ptr(), stream), "sfc order")
        idx = order.to(torch.int64)
        for name in ("pos", "vel", "orientation", "tag"):
            arr = getattr(self, name)
            arr[:n] = arr[:n].index_select(0, idx)
        return idx

    def positions_numpy(self):
        return self.pos.cpu().numpy()[:, :3].copy()
