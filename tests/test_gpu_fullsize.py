"""GPU, BASELINE.json's full sizes (C2 1 M, C3 4 M, C4 8 M, C5 16 M): size-independent
properties (Newton's third law over the full list, bit-reproducibility, mean row length) plus an
oracle check on a bounded sample of rows -- a prefix (`n_rows`) or, for C3-C5, rows drawn from
the whole system (helpers.oracle_compute_rows), which for C3 include colloid rows, i.e. the
warp-per-row LONGPASS kernel at its real 1,860-entry skew. Strict budgets (helpers.py)."""

import numpy as np
import pytest

import helpers
from oracle import oracle

pytestmark = pytest.mark.gpu


def test_c2_full_size_properties():
    """C2 at N = 1,000,000: (1) Newton's third law -- the full list holds both directions of every
    pair, so the net force vanishes; (2) the virial trace equals 1/2 sum r.F summed per pair, here
    checked as W_xx+W_yy+W_zz = 1/2 sum_i sum_j rsq*force_divr through symmetry W_ab = W_ba sums;
    (3) energies/forces of the first 20,000 rows match the oracle; (4) two evaluations are
    bit-identical (deterministic, no atomics)."""
    import torch

    import azplugins_b200 as az
    from azplugins_b200 import synth

    wl = synth.config2()
    assert wl.N == 1000000
    state = wl.make_state(dtype=np.float32)
    nl = az.nlist.Cell(buffer=synth.BUFFER)
    (pot,) = wl.make_potentials(nl)
    pot.attach(state).compute()
    F = pot._force.double()
    net = F[:, :3].sum(0).abs().max().item()
    scale = F[:, :3].abs().sum(0).max().item()
    assert net <= 1e-6 * scale
    n_bar = nl.n_neigh.double().mean().item()
    assert abs(n_bar - wl.n_bar) / wl.n_bar < 0.02
    first = pot._force.clone()
    pot.compute()
    assert torch.equal(first, pot._force)
    n_rows = 20000
    ref = helpers.oracle_compute(oracle.load("best", np.float32), state, pot, nl.to_numpy(), n_rows=n_rows)
    helpers.check_against_oracle(pot, ref, 4, n_rows=n_rows)


def test_c5_large_net_force_and_torque_balance():
    """C5 shape at N = 2,000,000: net force vanishes; rows 0..20k match the oracle."""
    import azplugins_b200 as az
    from azplugins_b200 import synth

    wl = synth.config5(N=2000000)
    state = wl.make_state(dtype=np.float32)
    nl = az.nlist.Cell(buffer=synth.BUFFER)
    (pot,) = wl.make_potentials(nl)
    pot.attach(state).compute(compute_virial=False)
    F = pot._force.double()
    assert F[:, :3].sum(0).abs().max().item() <= 1e-6 * F[:, :3].abs().sum(0).max().item()
    n_rows = 20000
    ref = helpers.oracle_compute(oracle.load("best", np.float32), state, pot, nl.to_numpy(),
                                 virial=False, n_rows=n_rows)
    helpers.check_against_oracle(pot, ref, 4, n_rows=n_rows, virial=False)


def _full_size(cfg, N):
    import torch

    import azplugins_b200 as az
    from azplugins_b200 import synth

    wl = synth.CONFIGS[cfg](N=N)
    state = wl.make_state(dtype=np.float32)
    nl = az.nlist.Cell(buffer=synth.BUFFER)
    pots = wl.make_potentials(nl)
    for pot in pots:
        pot.attach(state)
    torch.cuda.synchronize()
    return wl, state, nl, pots


def _net_force_vanishes(pot, tol=1e-6):
    F = pot._force.double()
    net = F[:, :3].sum(0).abs().max().item()
    scale = F[:, :3].abs().sum(0).max().item()
    assert net <= tol * scale, (net, scale)


def _sample(rng, N, m, extra=()):
    rows = np.unique(np.concatenate([rng.choice(N, m, replace=False), np.asarray(extra, dtype=np.int64)]))
    return rows


def test_c3_full_size_4m_colloid_rows_through_the_long_pass():
    """C3 at N = 4,000,000 (2,000 colloids, 1,860-entry rows): Colloid and Hertz each against the
    oracle on 150 colloid rows (second, warp-per-row pass) + 6,000 solvent rows + every
    solvent neighbour row of one colloid (colloid-solvent branch); net force vanishes;
    bit-reproducible although the long-row queue is filled with atomics."""
    import torch

    wl, state, nl, pots = _full_size("C3", 4000000)
    assert abs(wl.N - 4000000) < 40000
    typeid = wl.typeid
    colloids = np.nonzero(typeid == 1)[0]
    assert len(colloids) == 2000
    nl.compute(state)
    nn = nl.n_neigh.cpu().numpy()
    assert nn[colloids].min() > 512 and nl.n_max > 512  # these rows do take the second pass
    rng = np.random.default_rng(5)
    c0 = int(colloids[7])
    h0 = int(nl.head_list[c0].item())
    around = nl.nlist[h0:h0 + int(nn[c0])].cpu().numpy().astype(np.int64)[::4]
    rows = _sample(rng, wl.N, 6000, extra=np.concatenate([rng.choice(colloids, 150, replace=False),
                                                          [c0], around]))
    orc = oracle.load("best", np.float32)
    for pot in pots:
        pot.compute()
        first = pot._force.clone()
        _net_force_vanishes(pot)
        ref = helpers.oracle_compute_rows(orc, state, pot, nl, rows)
        rep = helpers.check_against_oracle(pot, ref, 4, rows=rows)
        assert set(rep["criterion"].values()) == {"strict"}
        print("C3 4M", type(pot).__name__, helpers.format_report(rep))
        pot.compute()
        assert torch.equal(first, pot._force)


def test_c4_full_size_8m_dpd_thermostat():
    """C4 at N = 8,000,000: thermostatted forces of 20,000 sampled rows against the oracle
    (same Philox stream); the pairwise drag and random forces are antisymmetric, so the net
    force vanishes over the full list (alpha_ij = alpha_ji at 8 M tags)."""
    wl, state, nl, (pot,) = _full_size("C4", 8000000)
    assert wl.N == 8000000
    pot.compute()
    _net_force_vanishes(pot, tol=2e-6)
    # (the mean row length of the 200^3 jittered lattice at r_list = 1.4 is 29.9, not the ideal-
    # gas 34.5 of wl.n_bar: lattice shells)
    assert 25.0 < nl.n_neigh.double().mean().item() < 40.0
    rows = _sample(np.random.default_rng(6), wl.N, 20000)
    ref = helpers.oracle_compute_rows(oracle.load("best", np.float32), state, pot, nl, rows)
    rep = helpers.check_against_oracle(pot, ref, 4, rows=rows)
    assert set(rep["criterion"].values()) == {"strict"}
    print("C4 8M", helpers.format_report(rep))


def test_c5_full_size_16m_forces_and_torques():
    """C5 at N = 16,000,000: forces, torques, energies and virials of 20,000 sampled rows against
    the oracle; net force vanishes."""
    wl, state, nl, (pot,) = _full_size("C5", 16000000)
    assert wl.N == 16000000
    pot.compute()
    _net_force_vanishes(pot)
    rows = _sample(np.random.default_rng(7), wl.N, 20000)
    ref = helpers.oracle_compute_rows(oracle.load("best", np.float32), state, pot, nl, rows)
    rep = helpers.check_against_oracle(pot, ref, 4, rows=rows)
    assert set(rep["criterion"].values()) == {"strict"}
    print("C5 16M", helpers.format_report(rep))
