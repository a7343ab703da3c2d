"""Row-range launches (``compute(rows=(lo, hi))``) and the streamed host delivery
(``compute_to_host``) give the same bits as one launch over all rows."""
import numpy as np
import pytest
import torch

import azplugins_b200 as az
from azplugins_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg,N", [("C2", 20000), ("C4", 24000), ("C5", 32768), ("C3", 60000)])
def test_row_ranges_equal_full_launch(cfg, N):
    wl = synth.CONFIGS[cfg](N=N)
    state = wl.make_state(dtype=np.float32, device="cuda:0")
    nl = az.nlist.Cell(buffer=synth.BUFFER)
    pots = wl.make_potentials(nl)
    for pot in pots:
        pot.attach(state)
        pot.kernel_parameters = (128, 2)
        pot.compute(compute_virial=True)
        torch.cuda.synchronize()
        f0, v0, t0 = pot._force.clone(), pot._virial.clone(), pot._torque.clone()
        pot._force.fill_(7.0)
        pot._virial.fill_(7.0)
        pot._torque.fill_(7.0)
        n = state.N
        cuts = [0, 1, n // 3 + 5, n // 3 + 5, n - 7, n]  # includes an empty range
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            pot.compute(compute_virial=True, rows=(lo, hi))
        torch.cuda.synchronize()
        if cfg == "C4":
            # DPD: summation order depends on the rows sharing a warp (deferred-accept queue)
            assert torch.allclose(pot._force, f0, rtol=0, atol=2e-4 * f0.abs().max().item())
            assert torch.allclose(pot._virial, v0, rtol=0, atol=2e-4 * v0.abs().max().item())
        else:
            assert torch.equal(pot._force, f0)
            assert torch.equal(pot._virial, v0)
            if pot.is_anisotropic:
                assert torch.equal(pot._torque, t0)
        with pytest.raises(ValueError):
            pot.compute(rows=(5, n + 1))


def test_compute_to_host_matches_device_results():
    wl = synth.config2(N=30000)
    state = wl.make_state(dtype=np.float32, device="cuda:0")
    nl = az.nlist.Cell(buffer=synth.BUFFER)
    (pot,) = wl.make_potentials(nl)
    pot.attach(state)
    pot.kernel_parameters = (128, 1)
    pot.compute(compute_virial=True)
    torch.cuda.synchronize()
    f0, v0 = pot._force.cpu(), pot._virial.cpu()
    hf = torch.empty_like(f0).pin_memory()
    hv = torch.empty_like(v0).pin_memory()
    for chunks in (1, 3, 4):
        hf.fill_(-1.0)
        hv.fill_(-1.0)
        pot.compute_to_host(hf, hv, chunks=chunks)
        assert torch.equal(hf, f0)
        assert torch.equal(hv, v0)


def test_nlist_capacity_reuse_gives_the_same_list():
    """A rebuild that reuses the previous row capacities (no count pass) yields the same rows as a
    from-scratch build; an overflowing row falls back to count + fill."""
    wl = synth.config2(N=30000)
    state = wl.make_state(dtype=np.float32, device="cuda:0")
    nl = az.nlist.Cell(buffer=synth.BUFFER)
    (pot,) = wl.make_potentials(nl)
    pot.attach(state)
    nl.build(state)
    assert nl.num_reused == 0
    ref = [t.clone() for t in (nl.n_neigh, nl.nlist, nl.head_list)]
    # small displacements: same capacities suffice
    g = torch.Generator(device="cuda").manual_seed(1)
    state.pos[:, :3] += 0.05 * (torch.rand(state.pos[:, :3].shape, device="cuda", generator=g) - 0.5)
    nl.build(state)
    assert nl.num_reused == 1
    reused = [t.clone() for t in (nl.n_neigh, nl.nlist, nl.head_list)]
    nl.reuse_capacity = False
    nl.build(state)
    assert nl.num_reused == 1
    nn = nl.n_neigh.long()
    assert torch.equal(reused[0], nl.n_neigh)
    # same valid entries row by row (capacities, hence heads, may differ)
    for r in (0, 1, 777, 29999):
        a = reused[1][int(reused[2][r]):int(reused[2][r]) + int(nn[r])]
        b = nl.nlist[int(nl.head_list[r]):int(nl.head_list[r]) + int(nn[r])]
        assert torch.equal(a, b)
    # overflow: squeeze many particles together so that some rows outgrow their capacity
    nl.reuse_capacity = True
    nl.build(state)
    reused_before = nl.num_reused
    state.pos[:2000, :3] = state.pos[:1, :3] + 0.3 * torch.rand((2000, 3), device="cuda", generator=g)
    nl.build(state)
    assert nl.num_reused == reused_before  # fell back to the count pass
    assert int(nl.n_neigh.max()) >= 1999
    pot.compute()
    torch.cuda.synchronize()
    del ref
