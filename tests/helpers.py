"""Shared test helpers: run a workload through the CUDA path (via the C ABI) and through the
CPU oracle on identical arrays and an identical neighbour list, and measure the differences with
the tolerances BASELINE.json states.

Error measures (all computed in float64):
  force / torque : max_i |F_i - F_i^ref| / max(|F_i^ref|, rms_k |F_k^ref|)
                   -- per-particle relative error, with the rms force as the floor so that
                   particles whose net force cancels to ~0 do not divide by ~0;
  energy         : |sum_i E_i - sum_i E_i^ref| / |sum_i E_i^ref|
  virial         : max_ab |sum_i W_i^ab - ref| / max_ab |ref^ab|
Budgets: fp32 1e-5 (force, torque), fp64 1e-10; energy and virial 1e-6 (fp32), 1e-10 (fp64).
"""

import numpy as np

FORCE_TOL = {4: 1e-5, 8: 1e-10}
TOTAL_TOL = {4: 1e-6, 8: 1e-10}


def per_particle_rel_err(a, ref):
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    n = np.linalg.norm(ref, axis=1)
    floor = np.sqrt(np.mean(n * n)) if n.size else 0.0
    if floor == 0.0:
        return float(np.abs(a - ref).max()) if a.size else 0.0
    return float((np.linalg.norm(a - ref, axis=1) / np.maximum(n, floor)).max())


def total_rel_err(a, ref):
    sa = float(np.asarray(a, dtype=np.float64).sum())
    sr = float(np.asarray(ref, dtype=np.float64).sum())
    if sr == 0.0:
        return abs(sa)
    return abs(sa - sr) / abs(sr)


def virial_rel_err(v, vref):
    """v: (N,6) product layout; vref: (6,N) oracle layout."""
    sv = np.asarray(v, dtype=np.float64).sum(axis=0)
    sr = np.asarray(vref, dtype=np.float64).sum(axis=1)
    den = np.abs(sr).max()
    if den == 0.0:
        return float(np.abs(sv).max())
    return float(np.abs(sv - sr).max() / den)


def oracle_tables(orc, state_types, pot):
    """param table / r_cut / r_on matrices for the oracle from a pair.* object."""
    nt = len(state_types)
    name = ORACLE_NAME[type(pot).__name__]
    pp = {}
    rc = np.zeros((nt, nt))
    ro = np.zeros((nt, nt))
    for i, a in enumerate(state_types):
        for j, b in enumerate(state_types):
            pp[(i, j)] = pot.params[(a, b)]
            rc[i, j] = pot.r_cut[(a, b)]
            ro[i, j] = pot.r_on[(a, b)]
    return name, orc.pack_table(name, nt, pp), rc, ro


ORACLE_NAME = {
    "PerturbedLennardJones": "PerturbedLennardJones",
    "ExpandedYukawa": "ExpandedYukawa",
    "Colloid": "Colloid",
    "Hertz": "Hertz",
    "DPDGeneralWeight": "DPDGeneralWeight",
    "DPDGeneralWeightConservative": "DPDGeneralWeight",
    "TwoPatchMorse": "TwoPatchMorse",
}


def oracle_compute(orc, state, pot, nl_arrays, virial=True, n_rows=None, half=False,
                   rint_image=False):
    """Run the matching oracle loop on the host copies of ``state``'s arrays.
    Returns dict(force (N,4), virial (6,N), torque (N,4) or None)."""
    from azplugins_b200 import pair

    n_neigh, nlist, head = nl_arrays
    pos = state.pos.cpu().numpy()
    name, table, rc, ro = oracle_tables(orc, state.types, pot)
    L = state.box.L
    tilt = (state.box.xy, state.box.xz, state.box.yz)
    per = [int(p) for p in state.box.periodic]
    N = state.N if n_rows is None else n_rows
    common = dict(ntypes=state.ntypes, virial=virial, tilt=tilt, periodic=per, half=half,
                  rint_image=rint_image, N=N)
    if isinstance(pot, pair.TwoPatchMorse):
        f, t, v = orc.aniso_forces(table, pos, state.orientation.cpu().numpy(), n_neigh, nlist,
                                   head, L, rc, mode=pot.mode, **common)
        return dict(force=f, torque=t, virial=v)
    if type(pot) is pair.DPDGeneralWeight:
        f, v = orc.dpd_forces(table, pos, state.vel.cpu().numpy(),
                              state.tag.cpu().numpy().view(np.uint32), n_neigh, nlist, head, L,
                              rc, state.seed, state.timestep, state.dt,
                              pot._kT(state.timestep), **common)
        return dict(force=f, torque=None, virial=v)
    f, v = orc.pair_forces(name, table, pos, n_neigh, nlist, head, L, rc, r_on=ro,
                           mode=pot.mode, **common)
    return dict(force=f, torque=None, virial=v)


def check_against_oracle(pot, ref, itemsize, n_rows=None, force_tol=None, total_tol=None,
                         virial=True):
    """Assert the product's read-outs match an oracle result within the stated budgets."""
    ftol = FORCE_TOL[itemsize] if force_tol is None else force_tol
    ttol = TOTAL_TOL[itemsize] if total_tol is None else total_tol
    sl = slice(0, n_rows)
    F = pot.forces[sl]
    E = pot.energies[sl]
    ef = per_particle_rel_err(F, ref["force"][sl, :3])
    ee = total_rel_err(E, ref["force"][sl, 3])
    report = dict(force=ef, energy=ee)
    assert ef <= ftol, "per-particle force rel err %.3e > %.1e" % (ef, ftol)
    assert ee <= ttol, "total energy rel err %.3e > %.1e" % (ee, ttol)
    if virial:
        ev = virial_rel_err(pot.virials[sl], ref["virial"][:, sl])
        report["virial"] = ev
        assert ev <= ttol, "total virial rel err %.3e > %.1e" % (ev, ttol)
    if ref.get("torque") is not None:
        et = per_particle_rel_err(pot.torques[sl], ref["torque"][sl, :3])
        report["torque"] = et
        assert et <= ftol, "per-particle torque rel err %.3e > %.1e" % (et, ftol)
    return report
