"""GPU: the fused two-potential pass (``azp_pair_forces_fused_*``):
Colloid + Hertz over one sweep of the list give the outputs of the two separate launches --
including the colloid rows that go through the warp-per-row long pass. Hertz is bit-identical;
Colloid to the order of summation: its separate launch defers the rare sphere-point /
sphere-sphere pairs of a row behind the common ones (pair_kernels.cuh, FormSplit), the fused
pass evaluates every pair in list order."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("virial", [False, True])
def test_fused_colloid_hertz_equals_separate_launches(dtype, virial):
    import torch

    import azplugins_b200 as az
    from azplugins_b200 import synth

    # 195 colloids on an 11.6-sigma lattice with jitter: 18 centre distances fall below the Hertz cutoff
    wl = synth.config3(N=120000, n_colloid=6500)
    state = wl.make_state(dtype=dtype)
    nl = az.nlist.Cell(buffer=synth.BUFFER)
    colloid, hertz = wl.make_potentials(nl)
    tol = 2e-6 if dtype == np.float32 else 1e-14

    def same(pot, got, want):
        if pot is hertz:
            return torch.equal(got, want)
        return float((got - want).abs().max()) <= tol * float(want.abs().max())

    for mode_c, mode_h in (("none", "none"), ("shift", "none")):
        colloid.mode, hertz.mode = mode_c, mode_h
        for shape in ((128, 1), (128, 4), (64, 32)):
            for p in (colloid, hertz):
                p.attach(state)
                p.kernel_parameters = shape
                p.compute(compute_virial=virial)
            want = [(p._force.clone(), p._virial.clone()) for p in (colloid, hertz)]
            assert float(hertz._force.abs().max()) > 0  # colloid-colloid contacts exist
            for p in (colloid, hertz):
                p._force.fill_(7.0)
                p._virial.fill_(7.0)
            fused = az.pair.FusedPair(hertz, colloid)  # either order
            fused.kernel_parameters = shape
            fused.compute(compute_virial=virial)
            for p, (f, w) in zip((colloid, hertz), want):
                assert same(p, p._force, f), (type(p).__name__, shape, mode_c)
                if virial:
                    assert same(p, p._virial, w), (type(p).__name__, shape, mode_c)
    # a row range (what compute_to_host and the slice scheduler use)
    lo, hi = 1000, 50000
    for p in (colloid, hertz):
        p._force.fill_(3.0)
    fused.compute(compute_virial=virial, rows=(lo, hi))
    for p, (f, w) in zip((colloid, hertz), want):
        assert same(p, p._force[lo:hi], f[lo:hi])
        assert bool((p._force[:lo] == 3.0).all()) and bool((p._force[hi:] == 3.0).all())


def test_fused_rejects_what_it_cannot_fuse():
    import azplugins_b200 as az
    from azplugins_b200 import synth

    wl = synth.config3(N=120000)
    state = wl.make_state()
    nl = az.nlist.Cell(buffer=synth.BUFFER)
    colloid, hertz = wl.make_potentials(nl)
    assert az.pair.FusedPair.can_fuse(colloid, hertz) and az.pair.FusedPair.can_fuse(hertz, colloid)
    hertz.mode = "xplor"
    assert not az.pair.FusedPair.can_fuse(colloid, hertz)
    with pytest.raises(ValueError):
        az.pair.FusedPair(colloid, hertz)
    hertz.mode = "none"
    other = az.pair.Hertz(nlist=az.nlist.Cell(buffer=0.4), default_r_cut=1.0)
    with pytest.raises(ValueError):
        az.pair.FusedPair(colloid, other)  # different neighbour lists
    with pytest.raises(ValueError):
        az.pair.FusedPair(hertz, hertz)  # no fused kernel for this pair
    del state
