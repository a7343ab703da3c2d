"""Synthetic workloads of BASELINE.json: jittered-lattice fluids in a cubic
periodic box, particles sorted along a Morton curve (what HOOMD's SFC sorter would do), seeds
fixed. Host-side numpy only; nothing here is on the timed path.

    C1  PerturbedLennardJones  N=32,000      rho=0.8  r_cut=3.0  1 type
    C2  ExpandedYukawa         N=1,000,000   rho=0.5  r_cut=3.5  2 types, shift, virial
    C3  Colloid + Hertz        N~4,000,000   2,000 colloids (a=5) in solvent rho_S=0.7
    C4  DPDGeneralWeight       N=8,000,000   rho=3    r_cut=1.0  thermostat
    C5  TwoPatchMorse          N=16,000,000  rho=0.6  r_cut=1.6  forces + torques
Each config accepts a smaller ``N`` (same density and parameters) for parity tests.
"""

import numpy as np

from .box import Box
from .state import State

BUFFER = 0.4  # the buffer every reference test uses (src/pytest/test_pair.py:318)


def _lattice_sites(N):
    """(fractional sites in [0,1)^3 of an n^3 cell lattice, n, basis size) with n^3*basis >= N."""
    for basis in ([(0, 0, 0)],
                  [(0, 0, 0), (0.5, 0.5, 0.5)],
                  [(0, 0, 0), (0.5, 0.5, 0), (0.5, 0, 0.5), (0, 0.5, 0.5)]):
        n = round((N / len(basis)) ** (1.0 / 3.0))
        if n > 0 and n ** 3 * len(basis) == N:
            return np.array(basis, dtype=np.float64), n
    n = int(np.ceil(N ** (1.0 / 3.0)))
    return np.array([(0, 0, 0)], dtype=np.float64), n


def jittered_lattice(N, rho, rng, jitter=0.25):
    """N points at number density rho: lattice sites + uniform jitter of +-jitter * nearest-
    neighbour distance / sqrt(3) per axis (so no two particles come closer than ~0.5 d_nn)."""
    basis, n = _lattice_sites(N)
    L = (N / rho) ** (1.0 / 3.0)
    a = L / n
    g = np.arange(n, dtype=np.float64)
    cells = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    sites = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a
    if sites.shape[0] > N:
        keep = rng.choice(sites.shape[0], size=N, replace=False)
        keep.sort()
        sites = sites[keep]
    if len(basis) == 1:
        dnn = a
    elif len(basis) == 2:
        dnn = a * np.sqrt(3.0) / 2.0
    else:
        dnn = a / np.sqrt(2.0)
    amp = jitter * dnn / np.sqrt(3.0)
    xyz = sites + rng.uniform(-amp, amp, size=sites.shape)
    xyz -= 0.5 * L
    xyz = (xyz + 0.5 * L) % L - 0.5 * L
    return xyz, L


def morton_order(xyz, L, cell=1.0):
    """Permutation sorting particles along a Morton (Z-order) curve of a grid of ~`cell` cells."""
    n = int(min(1024, max(1, np.floor(L / cell))))
    c = np.floor((xyz / L + 0.5) * n).astype(np.int64)
    c = np.clip(c, 0, n - 1)

    def spread(v):
        v = v & 0x3FF
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v

    code = spread(c[:, 0]) | (spread(c[:, 1]) << 1) | (spread(c[:, 2]) << 2)
    return np.argsort(code, kind="stable")


class Workload:
    """A synthetic configuration: arrays + the potentials (class name, params, r_cut, mode)."""

    def __init__(self, name, box, types, position, typeid, potentials, velocity=None,
                 orientation=None, tag=None, seed=0, timestep=0, dt=0.005, n_bar=None,
                 bytes_per_particle=None, compute_virial=False):
        self.name = name
        self.box = box
        self.types = types
        self.position = position
        self.typeid = typeid
        self.velocity = velocity
        self.orientation = orientation
        self.tag = tag
        self.potentials = potentials
        self.seed = seed
        self.timestep = timestep
        self.dt = dt
        self.n_bar = n_bar
        self.bytes_per_particle = bytes_per_particle
        self.compute_virial = compute_virial

    @property
    def N(self):
        return self.position.shape[0]

    def make_state(self, dtype=np.float32, device="cuda:0"):
        return State(self.box, self.types, self.position, typeid=self.typeid,
                     velocity=self.velocity, orientation=self.orientation, tag=self.tag,
                     dtype=dtype, device=device, seed=self.seed, timestep=self.timestep,
                     dt=self.dt)

    def make_potentials(self, nlist):
        """Instantiate the ``azplugins_b200.pair`` classes on ``nlist``."""
        from . import pair

        out = []
        for spec in self.potentials:
            cls = getattr(pair, spec["cls"])
            kwargs = dict(spec.get("kwargs", {}))
            pot = cls(nlist=nlist, default_r_cut=spec["default_r_cut"], **kwargs)
            for key, p in spec["params"].items():
                pot.params[key] = p
            for key, rc in spec.get("r_cut", {}).items():
                pot.r_cut[key] = rc
            out.append(pot)
        return out


def _finish(xyz, L, rng, extra=()):
    perm = morton_order(xyz, L)
    out = [xyz[perm]]
    for e in extra:
        out.append(None if e is None else e[perm])
    return perm, out


def _nbar(rho, r_list):
    return 4.0 / 3.0 * np.pi * r_list ** 3 * rho


def config1(N=32000, seed=20261):
    """C1: PerturbedLennardJones single-type fluid, rho=0.8, r_cut=3.0, mode none."""
    rng = np.random.default_rng(seed)
    xyz, L = jittered_lattice(N, 0.8, rng)
    perm, (xyz,) = _finish(xyz, L, rng)
    pots = [dict(cls="PerturbedLennardJones", default_r_cut=3.0, kwargs=dict(mode="none"),
                 params={("A", "A"): dict(epsilon=1.0, sigma=1.0, attraction_scale_factor=0.5)})]
    nb = _nbar(0.8, 3.0 + BUFFER)
    return Workload("C1-PerturbedLennardJones", Box.cube(L), ["A"], xyz,
                    np.zeros(len(xyz), dtype=np.uint32), pots, tag=perm.astype(np.uint32),
                    n_bar=nb, bytes_per_particle=4 + 8 + 16 + 16 + 4 * nb)


def config2(N=1000000, seed=20262):
    """C2: ExpandedYukawa, 2 types 50:50, rho=0.5, r_cut=3.5, shift, virial on."""
    rng = np.random.default_rng(seed)
    xyz, L = jittered_lattice(N, 0.5, rng)
    typeid = (rng.random(len(xyz)) < 0.5).astype(np.uint32)
    perm, (xyz, typeid) = _finish(xyz, L, rng, (typeid,))
    pots = [dict(cls="ExpandedYukawa", default_r_cut=3.5, kwargs=dict(mode="shift"),
                 params={("A", "A"): dict(epsilon=1.0, kappa=1.0, delta=0.0),
                         ("A", "B"): dict(epsilon=2.0, kappa=1.2, delta=0.15),
                         ("B", "B"): dict(epsilon=3.0, kappa=1.5, delta=0.3)})]
    nb = _nbar(0.5, 3.5 + BUFFER)
    return Workload("C2-ExpandedYukawa", Box.cube(L), ["A", "B"], xyz, typeid, pots,
                    tag=perm.astype(np.uint32), n_bar=nb,
                    bytes_per_particle=4 + 8 + 16 + 16 + 4 * nb + 24, compute_virial=True)


def config3(N=4000000, n_colloid=2000, seed=20263):
    """C3: colloids (radius 5) in explicit solvent; Colloid + Hertz over one list.
    The solvent is a rho_S=0.7 jittered lattice with a r < 5.9 hole around every colloid, so the
    particle count is close to, not exactly, N."""
    rng = np.random.default_rng(seed)
    a_c, hole, rho_s = 5.0, 5.9, 0.7
    scale = N / 4000000.0
    n_colloid = max(1, int(round(n_colloid * scale)))
    vol = (N - n_colloid) / rho_s + n_colloid * 4.0 / 3.0 * np.pi * hole ** 3
    L = vol ** (1.0 / 3.0)
    nc = int(np.ceil(n_colloid ** (1.0 / 3.0)))
    spacing = L / nc
    if spacing < 2 * a_c + 1.0:
        raise ValueError("colloids do not fit: increase N or reduce n_colloid")
    g = (np.arange(nc) + 0.5) * spacing - 0.5 * L
    csites = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    csites = csites[rng.choice(len(csites), n_colloid, replace=False)]
    # jitter colloids a little but keep centre distances >= 10.1 (contact at 10)
    max_j = max(0.0, 0.5 * (spacing - (2 * a_c + 0.1)))
    csites = csites + rng.uniform(-min(max_j, 1.0), min(max_j, 1.0), size=csites.shape)
    n_lat = int(round(rho_s * L ** 3))
    sxyz, _ = jittered_lattice(n_lat, n_lat / L ** 3, rng)
    keep = np.ones(len(sxyz), dtype=bool)
    # remove solvent inside the holes: the solvent is binned into cells at least one hole radius
    # wide, so a colloid only has to look at its 27 surrounding cells (same decisions as testing
    # every solvent particle against every colloid, which took minutes at N = 4 M)
    ncell = int(L // hole)
    if ncell >= 3:
        w = L / ncell
        cell3 = np.clip(np.floor((sxyz + 0.5 * L) / w).astype(np.int64), 0, ncell - 1)
        cid = (cell3[:, 2] * ncell + cell3[:, 1]) * ncell + cell3[:, 0]
        order = np.argsort(cid, kind="stable")
        start = np.searchsorted(cid[order], np.arange(ncell ** 3 + 1))
        off = np.array([-1, 0, 1])
        for c in csites:
            cc = np.clip(np.floor((c + 0.5 * L) / w).astype(np.int64), 0, ncell - 1)
            nx, ny, nz = ((cc[0] + off) % ncell), ((cc[1] + off) % ncell), ((cc[2] + off) % ncell)
            cells = ((nz[:, None, None] * ncell + ny[None, :, None]) * ncell + nx[None, None, :]).reshape(-1)
            idx = np.concatenate([order[start[k]:start[k + 1]] for k in np.unique(cells)])
            d = sxyz[idx] - c
            d -= L * np.round(d / L)
            keep[idx[(d ** 2).sum(axis=1) < hole ** 2]] = False
    else:
        for c in csites:
            d = sxyz - c
            d -= L * np.round(d / L)
            near = (np.abs(d) < hole).all(axis=1)
            idx = np.nonzero(near)[0]
            keep[idx[(d[idx] ** 2).sum(axis=1) < hole ** 2]] = False
    sxyz = sxyz[keep]
    xyz = np.concatenate([csites, sxyz])
    typeid = np.concatenate([np.ones(len(csites), dtype=np.uint32), np.zeros(len(sxyz), dtype=np.uint32)])
    perm, (xyz, typeid) = _finish(xyz, L, rng, (typeid,))
    pots = [dict(cls="Colloid", default_r_cut=3.0, kwargs=dict(mode="none"),
                 params={("S", "S"): dict(A=144.0, a_1=0.0, a_2=0.0, sigma=1.0),
                         ("S", "C"): dict(A=144.0, a_1=0.0, a_2=a_c, sigma=1.0),
                         ("C", "C"): dict(A=40.0, a_1=a_c, a_2=a_c, sigma=1.0)},
                 r_cut={("S", "C"): 9.0, ("C", "C"): 10.581}),
            dict(cls="Hertz", default_r_cut=0.0, kwargs=dict(mode="none"),
                 params={("S", "S"): dict(epsilon=0.0), ("S", "C"): dict(epsilon=0.0),
                         ("C", "C"): dict(epsilon=100.0)},
                 r_cut={("C", "C"): 10.581})]
    return Workload("C3-Colloid+Hertz", Box.cube(L), ["S", "C"], xyz, typeid, pots,
                    tag=perm.astype(np.uint32), n_bar=None, bytes_per_particle=None)


def config4(N=8000000, seed=20264, s=2.0):
    """C4: DPDGeneralWeight thermostat, rho=3, r_cut=1, kT=1, dt=0.01, seed 42, step 1000."""
    rng = np.random.default_rng(seed)
    xyz, L = jittered_lattice(N, 3.0, rng)
    vel = rng.standard_normal((len(xyz), 3))
    perm, (xyz, vel) = _finish(xyz, L, rng, (vel,))
    pots = [dict(cls="DPDGeneralWeight", default_r_cut=1.0, kwargs=dict(kT=1.0),
                 params={("A", "A"): dict(A=25.0, gamma=4.5, s=s)})]
    nb = _nbar(3.0, 1.0 + BUFFER)
    return Workload("C4-DPDGeneralWeight", Box.cube(L), ["A"], xyz,
                    np.zeros(len(xyz), dtype=np.uint32), pots, velocity=vel,
                    tag=perm.astype(np.uint32), seed=42, timestep=1000, dt=0.01, n_bar=nb,
                    bytes_per_particle=4 + 8 + 16 + 16 + 16 + 4 + 4 * nb)


def config5(N=16000000, seed=20265):
    """C5: TwoPatchMorse patchy particles, rho=0.6, r_cut=1.6, random orientations."""
    rng = np.random.default_rng(seed)
    xyz, L = jittered_lattice(N, 0.6, rng)
    q = rng.standard_normal((len(xyz), 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    perm, (xyz, q) = _finish(xyz, L, rng, (q,))
    pots = [dict(cls="TwoPatchMorse", default_r_cut=1.6, kwargs=dict(mode="none"),
                 params={("A", "A"): dict(M_d=1.8347, M_r=0.0302, r_eq=1.0043, omega=20.0,
                                          alpha=0.5, repulsion=True)})]
    nb = _nbar(0.6, 1.6 + BUFFER)
    return Workload("C5-TwoPatchMorse", Box.cube(L), ["A"], xyz,
                    np.zeros(len(xyz), dtype=np.uint32), pots, orientation=q,
                    tag=perm.astype(np.uint32), n_bar=nb,
                    bytes_per_particle=4 + 8 + 16 + 16 + 16 + 16 + 4 * nb)


CONFIGS = {"C1": config1, "C2": config2, "C3": config3, "C4": config4, "C5": config5}
