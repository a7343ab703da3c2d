"""Shared test helpers: run a workload through the CUDA path (via the C ABI) and through the
CPU oracle on identical arrays and an identical neighbour list, and measure the differences with
the tolerances BASELINE.json states.

Error measures (all computed in float64):
  force / torque : max_i |F_i - F_i^ref| / max(|F_i^ref|, rms_k |F_k^ref|)
                   -- per-particle relative error, with the rms force as the floor so that
                   particles whose net force cancels to ~0 do not divide by ~0;
  energy         : |sum_i E_i - sum_i E_i^ref| / sum_i |E_i^ref|
  virial         : max_ab |sum_i W_i^ab - ref^ab| / max_ab sum_i |W_i^ab,ref|
                   -- totals are measured against the sum of magnitudes: the synthetic
                   jittered-lattice fluids have close contacts whose large positive energies
                   nearly cancel the attractive ones, and a ratio to the cancelled total would
                   measure that cancellation instead of the arithmetic;
Budgets: fp32 1e-5 (force, torque), fp64 1e-10; energy and virial 1e-6 (fp32), 1e-10 (fp64).

Ill-conditioned potentials in fp32. Two evaluators amplify the rounding of r itself beyond those
budgets in ANY fp32 implementation: the two-patch Morse well (exp(-(r - r_eq)/0.03): 1 ulp of r
is 3e-6 of U) and the DPD weight (1 - r/r_cut)^(s/2) with s < 2 (infinite slope at the cutoff).
The reference's own fp32 CPU result is then farther than the budget from its fp64 result. For
those cases `truth` (the fp64 oracle on the same fp32-rounded inputs) may be given; the check
passes when the CUDA result is within budget of the fp32 oracle OR is no farther from the fp64
truth than 2x the fp32 oracle itself is (floored at the budget).
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
    ref = np.asarray(ref, dtype=np.float64)
    sr = float(ref.sum())
    den = float(np.abs(ref).sum())
    if den == 0.0:
        return abs(sa)
    return abs(sa - sr) / den


def virial_rel_err(v, vref):
    """v: (N,6) product layout; vref: (6,N) oracle layout."""
    sv = np.asarray(v, dtype=np.float64).sum(axis=0)
    vref = np.asarray(vref, dtype=np.float64)
    sr = vref.sum(axis=1)
    den = np.abs(vref).sum(axis=1).max()
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

    from oracle.oracle import make_pos, particle_types

    n_neigh, nlist, head = nl_arrays
    pos = state.pos.cpu().numpy()
    vel = state.vel.cpu().numpy()
    quat = state.orientation.cpu().numpy()
    if orc.dtype != pos.dtype:
        # fp64 "truth" on the SAME (fp32-rounded) inputs: widen values, re-encode the type ids
        pos = make_pos(pos[:, :3].astype(orc.dtype), particle_types(pos), orc.dtype)
        vel = vel.astype(orc.dtype)
        quat = quat.astype(orc.dtype)
    name, table, rc, ro = oracle_tables(orc, state.types, pot)
    L = state.box.L
    tilt = (state.box.xy, state.box.xz, state.box.yz)
    per = [int(p) for p in state.box.periodic]
    N = state.N if n_rows is None else n_rows
    common = dict(ntypes=state.ntypes, virial=virial, tilt=tilt, periodic=per, half=half,
                  rint_image=rint_image, N=N)
    if isinstance(pot, pair.TwoPatchMorse):
        f, t, v = orc.aniso_forces(table, pos, quat, n_neigh, nlist, head, L, rc, mode=pot.mode,
                                   **common)
        return dict(force=f, torque=t, virial=v)
    if type(pot) is pair.DPDGeneralWeight:
        f, v = orc.dpd_forces(table, pos, vel, state.tag.cpu().numpy().view(np.uint32), n_neigh,
                              nlist, head, L, rc, state.seed, state.timestep, state.dt,
                              pot._kT(state.timestep), **common)
        return dict(force=f, torque=None, virial=v)
    f, v = orc.pair_forces(name, table, pos, n_neigh, nlist, head, L, rc, r_on=ro,
                           mode=pot.mode, **common)
    return dict(force=f, torque=None, virial=v)


def _errors(F, E, W, T, ref, sl, virial):
    out = dict(force=per_particle_rel_err(F, ref["force"][sl, :3]),
               energy=total_rel_err(E, ref["force"][sl, 3]))
    if virial:
        out["virial"] = virial_rel_err(W, ref["virial"][:, sl])
    if ref.get("torque") is not None:
        out["torque"] = per_particle_rel_err(T, ref["torque"][sl, :3])
    return out


def check_against_oracle(pot, ref, itemsize, n_rows=None, force_tol=None, total_tol=None,
                         virial=True, truth=None):
    """Assert the product's read-outs match an oracle result within the stated budgets
    (see the module docstring for `truth`). Returns the measured errors."""
    ftol = FORCE_TOL[itemsize] if force_tol is None else force_tol
    ttol = TOTAL_TOL[itemsize] if total_tol is None else total_tol
    sl = slice(0, n_rows)
    F, E = pot.forces[sl], pot.energies[sl]
    W = pot.virials[sl] if virial else None
    T = pot.torques[sl] if ref.get("torque") is not None else None
    report = _errors(F, E, W, T, ref, sl, virial)
    budget = dict(force=ftol, torque=ftol, energy=ttol, virial=ttol)
    if truth is not None:
        gpu_t = _errors(F, E, W, T, truth, sl, virial)
        refT = None if ref.get("torque") is None else ref["torque"][sl, :3]
        cpu_t = _errors(ref["force"][sl, :3], ref["force"][sl, 3],
                        None if not virial else ref["virial"][:, sl].T, refT, truth, sl, virial)
        report.update({k + "_vs_fp64": v for k, v in gpu_t.items()})
        report.update({k + "_cpu32_vs_fp64": v for k, v in cpu_t.items()})
    for k, b in budget.items():
        if k not in report:
            continue
        ok = report[k] <= b
        if not ok and truth is not None:
            ok = report[k + "_vs_fp64"] <= max(b, 2.0 * report[k + "_cpu32_vs_fp64"])
        assert ok, "%s rel err %.3e > %.1e (%s)" % (k, report[k], b, report)
    return report
