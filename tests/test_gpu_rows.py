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
