"""CPU: the row-sample layout used by the full-size GPU parity tests
(helpers.oracle_compute_rows: sampled particles first, the whole system behind them as ghosts,
list entries shifted) reproduces the rows of a whole-system oracle run bit for bit."""

import numpy as np
import pytest

import helpers
from oracle import oracle


@pytest.mark.parametrize("cfg", ["C2", "C4", "C5"])
def test_sampled_rows_equal_whole_system_rows(cfg):
    import azplugins_b200 as az
    from azplugins_b200 import synth

    wl = synth.CONFIGS[cfg](N=4096)
    state = wl.make_state(dtype=np.float32, device="cpu")
    orc = oracle.load("port", np.float32)
    nl_dummy = az.nlist.NeighborList(buffer=synth.BUFFER)
    (pot,) = wl.make_potentials(nl_dummy)
    r_list = pot._r_cut_matrix(state) + synth.BUFFER
    arrays = orc.build_nlist(state.pos.numpy(), state.box.L, r_list, ntypes=state.ntypes)
    whole = helpers.oracle_compute(orc, state, pot, arrays)
    nl = az.nlist.NeighborList.from_arrays(*arrays, device="cpu")
    rows = np.unique(np.random.default_rng(1).choice(wl.N, 300, replace=False))
    part = helpers.oracle_compute_rows(orc, state, pot, nl, rows)
    assert np.array_equal(part["force"], whole["force"][rows])
    assert np.array_equal(part["virial"], whole["virial"][:, rows])
    if whole["torque"] is not None:
        assert np.array_equal(part["torque"], whole["torque"][rows])
