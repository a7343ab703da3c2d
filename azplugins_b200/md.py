"""A minimal ``hoomd.md.Integrator`` / ``md.methods.ConstantVolume`` work-alike: velocity Verlet
around the force computes of this package, so that the configurations run as real MD loops
(SURVEY.md 8(f) rank 4). The per-particle updates are the library's streaming kernels
``azp_nve_step_one/two`` (``include/azp_b200.h``); the neighbour list is rebuilt by the forces'
own ``nlist`` when a particle has moved more than half the buffer. CUDA only.

Not HOOMD's code path (HOOMD is not in the reference tree): the scheme is restated, its parity
with HOOMD is unpinned; ``tests/test_md.py`` checks it against a numpy restatement bit for bit and
through energy / momentum conservation.
"""

import ctypes

import torch

from . import _lib


class ConstantVolume:
    """``hoomd.md.methods.ConstantVolume(filter=All)`` without a thermostat: NVE."""


class Integrator:
    def __init__(self, dt, forces=None, methods=None):
        self.dt = float(dt)
        self.forces = list(forces or [])
        self.methods = list(methods or [ConstantVolume()])
        if len(self.methods) != 1 or not isinstance(self.methods[0], ConstantVolume):
            raise ValueError("only one ConstantVolume (NVE) method is supported")
        self._state = None

    def attach(self, state):
        if state.device.type != "cuda":
            raise _lib.AzpError("Integrator runs on CUDA devices only (no CPU fallback)")
        if any(t != 0.0 for t in (state.box.xy, state.box.xz, state.box.yz)):
            raise ValueError("orthorhombic boxes only")
        if not 1 <= len(self.forces) <= _lib.MD_MAX_FORCES:
            raise ValueError("need 1..%d forces" % _lib.MD_MAX_FORCES)
        self._state = state
        n = state.N
        self.accel = torch.zeros((n, 4), dtype=state.torch_dtype, device=state.device)
        self.net_force = torch.zeros((n, 4), dtype=state.torch_dtype, device=state.device)
        self.image = torch.zeros((n, 3), dtype=torch.int32, device=state.device)
        for f in self.forces:
            if f._state is not state:
                f.attach(state)
        self._prepared = False
        return self

    def _args(self):
        st = self._state
        a = _lib.AzpMdArgs()
        a.d_pos = st.pos.data_ptr()
        a.d_vel = st.vel.data_ptr()
        a.d_accel = self.accel.data_ptr()
        a.d_image = self.image.data_ptr()
        a.d_net_force = self.net_force.data_ptr()
        for k, f in enumerate(self.forces):
            a.d_forces[k] = f._force.data_ptr()
        a.n_forces = len(self.forces)
        a.N = st.N
        a.box = st.box.to_c()
        a.dt = self.dt
        return a

    def _call(self, name, args):
        st = self._state
        bits = 8 * st.dtype.itemsize
        with torch.cuda.device(st.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            rc = getattr(_lib.lib, "%s_f%d" % (name, bits))(ctypes.byref(args), stream)
        _lib.check(rc, name)

    def _compute_forces(self, compute_virial):
        from . import pair

        ts = self._state.timestep
        for f in self.forces:
            if isinstance(f, pair.Pair):
                f.compute(timestep=ts, compute_virial=compute_virial)
            else:  # one-body potentials (external, wall) always fill their virial array
                f.compute(timestep=ts)

    def run(self, steps, compute_virial=False):
        """Advance ``steps`` time steps (``sim.run(steps)``)."""
        st = self._state
        if st is None:
            raise RuntimeError("integrator is not attached to a State")
        if not self._prepared:
            # accelerations at the initial positions (HOOMD's prepRun); dt = 0 leaves v untouched
            self._compute_forces(compute_virial)
            a = self._args()
            a.dt = 0.0
            self._call("azp_nve_step_two", a)
            self._prepared = True
        a = self._args()
        for _ in range(int(steps)):
            self._call("azp_nve_step_one", a)
            st.timestep += 1
            self._compute_forces(compute_virial)
            self._call("azp_nve_step_two", a)
        return self

    # ---- thermodynamic read-outs (ComputeThermo's quantities) ------------------------------
    def kinetic_energy(self):
        v = self._state.vel.double()
        return float((0.5 * v[:, 3] * (v[:, :3] ** 2).sum(dim=1)).sum().item())

    def potential_energy(self):
        return float(self.net_force[:, 3].double().sum().item())

    def momentum(self):
        v = self._state.vel.double()
        return (v[:, 3:4] * v[:, :3]).sum(dim=0).cpu().numpy()
