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

Criterion. Every check reports, per quantity, WHICH criterion it passed on:
  "strict"      -- within the BASELINE.json budget of the oracle in the SAME precision;
  "fp64-escape" -- outside that budget, but no farther from the fp64 oracle (run on the same
                   fp32-rounded inputs) than 2x the fp32 oracle itself is (floored at the budget).
The escape is a builder-defined relaxation and is OFF unless a test passes ``truth=`` -- which
only the cases listed in ``ESCAPE_ALLOWED`` (and named in DESIGN.md section 5) may do. Round 2
made the two ill-conditioned fp32 potentials (two-patch Morse well, DPD weight with s < 2) round
their stiff intermediate results exactly where the reference's host code rounds them
(csrc/azp_core.cuh, namespace ref), so the list is empty: every case must pass "strict".
``report["criterion"]`` maps quantity -> criterion; smoke() prints it.
"""

import numpy as np

FORCE_TOL = {4: 1e-5, 8: 1e-10}
TOTAL_TOL = {4: 1e-6, 8: 1e-10}

# (config, potential class) pairs that may pass on the fp64-escape criterion. Empty on purpose.
ESCAPE_ALLOWED = frozenset()


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


def sample_rows_arrays(state, nl, rows):
    """A bounded oracle problem for an ARBITRARY set of rows of a large system (full-size parity,
    tests/test_gpu_fullsize.py). The oracle evaluates rows [0, m) of whatever arrays it is given,
    so the sample is laid out as a system of m "local" particles (copies of the sampled
    particles, in sample order) followed by the whole original particle array as "ghosts": the
    sampled rows' neighbour lists are copied out of the device list and their entries shifted by
    m. Same particles, same neighbours in the same order -> the same per-row arithmetic.
    Returns (state-like dict of host arrays, (n_neigh, nlist, head_list), m)."""
    import torch

    rows_t = torch.as_tensor(np.asarray(rows, dtype=np.int64), device=nl.n_neigh.device)
    m = int(rows_t.numel())
    nn = nl.n_neigh[rows_t].to(torch.int64)
    head = nl.head_list[rows_t]
    start = torch.cumsum(nn, 0) - nn
    total = int(nn.sum().item())
    k = torch.arange(total, device=nn.device) - torch.repeat_interleave(start, nn)
    src = torch.repeat_interleave(head, nn) + k
    entries = nl.nlist[src].to(torch.int64) + m
    arrays = (nn.cpu().numpy().astype(np.uint32), entries.cpu().numpy().astype(np.uint32),
              start.cpu().numpy().astype(np.uint64))

    def stacked(t):
        return torch.cat([t[rows_t], t[: state.pos.shape[0]]]).cpu().numpy()

    host = dict(pos=stacked(state.pos), vel=stacked(state.vel), orientation=stacked(state.orientation),
                tag=stacked(state.tag).view(np.uint32))
    return host, arrays, m


class _HostState:
    """Just enough of azplugins_b200.State for oracle_compute, over host arrays."""

    def __init__(self, state, host):
        import torch

        self.types, self.ntypes, self.box = state.types, state.ntypes, state.box
        self.seed, self.timestep, self.dt = state.seed, state.timestep, state.dt
        self.pos = torch.from_numpy(host["pos"])
        self.vel = torch.from_numpy(host["vel"])
        self.orientation = torch.from_numpy(host["orientation"])
        self.tag = torch.from_numpy(host["tag"].view(np.int32))
        self.N = self.pos.shape[0]


def oracle_compute_rows(orc, state, pot, nl, rows, virial=True):
    """Oracle result for the rows ``rows`` of a large system (see sample_rows_arrays): dict with
    force (m,4), virial (6,m), torque (m,4) or None, in sample order."""
    host, arrays, m = sample_rows_arrays(state, nl, rows)
    out = oracle_compute(orc, _HostState(state, host), pot, arrays, virial=virial, n_rows=m)
    out["force"] = out["force"][:m]
    if out.get("torque") is not None:
        out["torque"] = out["torque"][:m]
    if out.get("virial") is not None:
        out["virial"] = out["virial"][:, :m]
    return out


def _errors(F, E, W, T, ref, sl, virial):
    out = dict(force=per_particle_rel_err(F, ref["force"][sl, :3]),
               energy=total_rel_err(E, ref["force"][sl, 3]))
    if virial:
        out["virial"] = virial_rel_err(W, ref["virial"][:, sl])
    if ref.get("torque") is not None:
        out["torque"] = per_particle_rel_err(T, ref["torque"][sl, :3])
    return out


def check_against_oracle(pot, ref, itemsize, n_rows=None, force_tol=None, total_tol=None,
                         virial=True, truth=None, rows=None):
    """Assert the product's read-outs match an oracle result within the stated budgets
    (see the module docstring for `truth`). ``rows``: the product rows that correspond to the
    oracle's rows [0, len(rows)) (oracle_compute_rows). Returns the measured errors."""
    ftol = FORCE_TOL[itemsize] if force_tol is None else force_tol
    ttol = TOTAL_TOL[itemsize] if total_tol is None else total_tol
    sl = slice(0, n_rows)
    if rows is not None:
        import torch

        rt = torch.as_tensor(np.asarray(rows, dtype=np.int64), device=pot._force.device)
        fsel = pot._force[rt].cpu().numpy()
        F, E = fsel[:, :3], fsel[:, 3]
        W = pot._virial[:, rt].cpu().numpy().T if virial else None
        T = pot._torque[rt].cpu().numpy()[:, :3] if ref.get("torque") is not None else None
        sl = slice(0, len(rows))
    else:
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
    criterion = {}
    for k, b in budget.items():
        if k not in report:
            continue
        ok = report[k] <= b
        criterion[k] = "strict"
        if not ok and truth is not None:
            ok = report[k + "_vs_fp64"] <= max(b, 2.0 * report[k + "_cpu32_vs_fp64"])
            criterion[k] = "fp64-escape"
        assert ok, "%s rel err %.3e > %.1e (%s)" % (k, report[k], b, report)
    report["criterion"] = criterion
    return report


def format_report(report):
    """One line: quantity=error[criterion] ... (what smoke() and the tests print)."""
    crit = report.get("criterion", {})
    return " ".join("%s=%.2e[%s]" % (k, report[k], crit[k]) for k in ("force", "torque", "energy", "virial")
                    if k in crit)
