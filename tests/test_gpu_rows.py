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


def _brute_force_rows(xyz, typeid, box, r_list, rows):
    """Neighbour sets of `rows` by brute force in float64 with HOOMD's sequential minimum image."""
    L = np.array(box.L)
    out = []
    for i in rows:
        d = xyz[i] - xyz
        if box.periodic[2]:
            img = np.rint(d[:, 2] / L[2])
            d[:, 2] -= L[2] * img
            d[:, 1] -= L[2] * box.yz * img
            d[:, 0] -= L[2] * box.xz * img
        if box.periodic[1]:
            img = np.rint(d[:, 1] / L[1])
            d[:, 1] -= L[1] * img
            d[:, 0] -= L[1] * box.xy * img
        if box.periodic[0]:
            img = np.rint(d[:, 0] / L[0])
            d[:, 0] -= L[0] * img
        rsq = (d * d).sum(axis=1)
        ok = rsq < r_list[typeid[i], typeid] ** 2
        ok[i] = False
        out.append(np.nonzero(ok)[0])
    return out


@pytest.mark.parametrize("tilt,periodic", [((0.0, 0.0, 0.0), (True, True, True)),
                                           ((0.3, -0.2, 0.15), (True, True, True)),
                                           ((0.0, 0.0, 0.0), (True, True, False)),
                                           ((0.25, 0.0, 0.0), (True, False, True))])
@pytest.mark.parametrize("tpr", [0, 1, 4, 32])
def test_nlist_builder_triclinic_and_walls_against_brute_force(tilt, periodic, tpr):
    """The cell-list builder (cell-sorted positions, image numbers from the stencil instead of
    rint) against brute force: orthorhombic and triclinic boxes, non-periodic axes (particles up
    to the faces), two types with different r_list, any lanes-per-row setting; every setting gives
    the same rows in the same order."""
    import azplugins_b200 as az

    rng = np.random.default_rng(31)
    N = 6000
    box = az.Box(14.0, 12.0, 16.0, xy=tilt[0], xz=tilt[1], yz=tilt[2], periodic=periodic)
    f = rng.uniform(-0.5, 0.5, (N, 3))
    xyz = np.empty_like(f)
    xyz[:, 2] = f[:, 2] * box.Lz
    xyz[:, 1] = f[:, 1] * box.Ly + box.yz * xyz[:, 2]
    xyz[:, 0] = f[:, 0] * box.Lx + box.xy * xyz[:, 1] + (box.xz - box.xy * box.yz) * xyz[:, 2]
    typeid = rng.integers(0, 2, N)
    state = az.State(box, ["A", "B"], xyz, typeid=typeid, dtype=np.float64)
    nl = az.nlist.Cell(buffer=0.4, threads_per_row=tpr)
    plj = az.pair.PerturbedLennardJones(nlist=nl, default_r_cut=2.0)
    plj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0, attraction_scale_factor=0.5)
    plj.params[("A", "B")] = dict(epsilon=1.0, sigma=1.0, attraction_scale_factor=0.5)
    plj.params[("B", "B")] = dict(epsilon=1.0, sigma=1.0, attraction_scale_factor=0.5)
    plj.r_cut[("B", "B")] = 2.9
    plj.attach(state)
    nl.compute(state)
    nn, lst, head = nl.to_numpy()
    r_list = nl.r_cut_matrix(state) + 0.4
    rows = rng.choice(N, 300, replace=False)
    pos64 = state.pos.cpu().numpy()[:, :3].astype(np.float64)
    for i, want in zip(rows, _brute_force_rows(pos64, typeid, box, r_list, rows)):
        got = lst[head[i]:head[i] + nn[i]]
        assert np.array_equal(np.sort(got), want), i
    ref = az.nlist.Cell(buffer=0.4, threads_per_row=8)
    plj2 = az.pair.PerturbedLennardJones(nlist=ref, default_r_cut=2.0)
    plj2.params[("A", "A")] = plj2.params[("A", "B")] = plj2.params[("B", "B")] = dict(
        epsilon=1.0, sigma=1.0, attraction_scale_factor=0.5)
    plj2.r_cut[("B", "B")] = 2.9
    plj2.attach(state)
    ref.compute(state)
    rn, rl, rh = ref.to_numpy()
    assert np.array_equal(nn, rn) and np.array_equal(head, rh)
    for i in rows:
        assert np.array_equal(lst[head[i]:head[i] + nn[i]], rl[rh[i]:rh[i] + rn[i]])


def test_device_sfc_sort_matches_host_morton_order():
    """azp_sfc_order (State.sfc_sort): the device Morton order equals the host one of synth.py on
    a 1024^3 grid, all per-particle arrays are permuted together, and forces computed after the
    sort are the permuted forces."""
    import azplugins_b200 as az
    from azplugins_b200 import synth

    rng = np.random.default_rng(8)
    N = 20000
    xyz, L = synth.jittered_lattice(N, 0.8, rng)
    shuffle = rng.permutation(N)
    xyz = xyz[shuffle]
    state = az.State(az.Box.cube(L), ["A"], xyz, velocity=rng.standard_normal((N, 3)), dtype=np.float32)
    pos_before = state.pos.cpu().numpy()
    nl = az.nlist.Cell(buffer=0.4)
    plj = az.pair.PerturbedLennardJones(nlist=nl, default_r_cut=3.0)
    plj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0, attraction_scale_factor=0.5)
    f_before = plj.attach(state).compute()._force.cpu().numpy().copy()
    perm = state.sfc_sort().cpu().numpy()
    assert np.array_equal(np.sort(perm), np.arange(N))
    assert np.array_equal(pos_before[perm], state.pos.cpu().numpy())
    assert np.array_equal(state.tag.cpu().numpy(), perm.astype(np.int32))

    def keys(p, real):
        # the kernel's key: 1024^3 grid over the fractional coordinates, 10 bits per axis interleaved
        f = p[:, :3].astype(real) * real(1.0 / L) + real(0.5)
        f = f - np.floor(f)
        c = np.clip(np.floor(f * real(1024)).astype(np.int64), 0, 1023)
        k = np.zeros(len(p), dtype=np.int64)
        for d in range(3):
            v = c[:, d] & 0x3FF
            v = (v | (v << 16)) & 0x030000FF
            v = (v | (v << 8)) & 0x0300F00F
            v = (v | (v << 4)) & 0x030C30C3
            v = (v | (v << 2)) & 0x09249249
            k |= v << d
        return k

    # sortedness under the kernel's own fp32 key: at most the few particles whose cell flips with
    # the rounding of x * (1/L) may be out of place; ties keep the original order (stable sort)
    k32 = keys(pos_before, np.float32)[perm]
    assert (np.diff(k32) < 0).mean() < 2e-3
    tie = np.diff(k32) == 0
    assert np.all(np.diff(perm)[tie] > 0)
    # and it is the curve of synth.morton_order (float64 cells): same order up to those flips
    k64 = keys(pos_before.astype(np.float64), np.float64)[perm]
    assert (np.diff(k64) < 0).mean() < 2e-3
    nl2 = az.nlist.Cell(buffer=0.4)
    plj2 = az.pair.PerturbedLennardJones(nlist=nl2, default_r_cut=3.0)
    plj2.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0, attraction_scale_factor=0.5)
    f_after = plj2.attach(state).compute()._force.cpu().numpy()
    scale = np.abs(f_before[:, :3]).max()
    assert np.abs(f_after[:, :3] - f_before[perm, :3]).max() < 2e-5 * scale


def test_cached_launch_arguments_follow_every_change():
    """pair.Pair.compute() reuses the argument struct of the previous all-rows launch while
    nothing it points at has changed (pair.py, _args_all_rows): positions updated in place,
    a new launch shape, a new mode, new parameters and a rebuilt list must all reach the kernel.
    The check is a second potential object on the SAME neighbour list, set up from scratch each
    time (full argument path), evaluated first so that both see the same list."""
    import torch

    import azplugins_b200 as az
    from azplugins_b200 import synth

    wl = synth.config2(N=20000)
    state = wl.make_state(dtype=np.float32)
    nl = az.nlist.Cell(buffer=synth.BUFFER)
    (pot,) = wl.make_potentials(nl)
    pot.attach(state)
    pot.compute(compute_virial=True)

    def check(**kw):
        (p2,) = wl.make_potentials(nl)
        p2.mode = pot.mode
        for key in list(pot.params.keys()):
            p2.params[key] = pot.params[key]
        p2.attach(state)
        p2.kernel_parameters = pot.kernel_parameters
        p2.compute(compute_virial=True, **kw)
        pot.compute(compute_virial=True, **kw)  # cached struct whenever nothing changed
        assert torch.equal(pot._force, p2._force) and torch.equal(pot._virial, p2._virial)
        assert float(pot._force.abs().max()) > 0
        return pot._force.clone()

    f0 = check()
    f1 = check()
    assert torch.equal(f0, f1)
    # positions moved in place (same pointer)
    state.pos[:, 0] += 0.01 * torch.sin(torch.arange(state.N, device=state.pos.device, dtype=torch.float32))
    f2 = check()
    assert not torch.equal(f2, f1)
    pot.kernel_parameters = (64, 4)
    check()
    pot.mode = "none"
    f3 = check()
    pot.params[("A", "B")] = dict(epsilon=2.5, kappa=1.1, delta=0.1)
    f4 = check()
    assert not torch.equal(f4[:, :3], f3[:, :3])
    check(timestep=17)


@pytest.mark.parametrize("cfg,N", [("C2", 27000), ("C4", 24000)])
def test_builder_row_range_equals_the_rows_of_a_full_build(cfg, N):
    """``Cell.build(state, rows=(lo, hi))`` -- what a rank of the slice scheduler builds: the rows
    of particles [lo, hi) searched among ALL particles -- gives exactly rows lo..hi-1 of the full
    build (fine-grid and 27-cell sweep alike; n_neigh / head_list indexed by row - lo)."""
    import azplugins_b200 as az
    from azplugins_b200 import synth

    wl = synth.CONFIGS[cfg](N=N)
    state = wl.make_state(dtype=np.float32)
    full = az.nlist.Cell(buffer=synth.BUFFER)
    wl.make_potentials(full)
    full.build(state)
    fn, fl, fh = full.to_numpy()
    for lo, hi in ((0, 1000), (5000, 5001), (7777, state.N)):
        part = az.nlist.Cell(buffer=synth.BUFFER)
        wl.make_potentials(part)
        part.build(state, rows=(lo, hi))
        pn, pl, ph = part.to_numpy()
        assert len(pn) == hi - lo and np.array_equal(pn, fn[lo:hi])
        for r in range(0, hi - lo, max(1, (hi - lo) // 200)):
            a = pl[ph[r]:ph[r] + pn[r]]
            b = fl[fh[lo + r]:fh[lo + r] + fn[lo + r]]
            assert np.array_equal(a, b), (lo, hi, r)


@pytest.mark.parametrize("L,coarse", [(9.0, False), (30.0, False), (30.0, True)])
def test_builder_with_particles_on_and_outside_the_box_faces(L, coarse, monkeypatch):
    """Particles exactly on the upper box face (fp32 rounding puts one there about once per 10 M
    particles) or one box length outside are binned and swept with their in-box image: the list
    must be the brute-force minimum-image one. L = 9: fine grid at its smallest (6 half-width
    cells per axis), L = 30: 20 cells per axis; coarse: the 27-cell sweep (AZP_NLIST_COARSE)."""
    if coarse:
        monkeypatch.setenv("AZP_NLIST_COARSE", "1")
    import azplugins_b200 as az

    rng = np.random.default_rng(5)
    n = 1500 if L < 10 else 4000
    xyz = rng.uniform(-0.5, 0.5, size=(n, 3)) * L
    xyz[:30] = 0.5 * np.float64(np.float32(L)) * rng.choice([-1.0, 1.0], size=(30, 3))  # on faces / corners
    xyz[30:60] = np.where(rng.random((30, 3)) < 0.5, 0.5 * np.float64(np.float32(L)), xyz[30:60])
    xyz[60:90, 1] += L   # one box length outside
    xyz[90:120, 2] -= L
    box = az.Box.cube(L)
    # State validates nothing about the box faces; build it from raw arrays
    state = az.State(box, ["A"], xyz, dtype=np.float32)
    pos = state.pos.cpu().numpy()[:, :3].astype(np.float64)
    nl = az.nlist.Cell(buffer=0.4)
    pot = az.pair.PerturbedLennardJones(nlist=nl, default_r_cut=2.5)
    pot.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0, attraction_scale_factor=0.5)
    pot.attach(state)
    nl.build(state)
    nn, lst, head = nl.to_numpy()
    r_list = 2.9
    missing = extra = 0
    for i in range(n):
        d = pos[i] - pos
        d -= L * np.round(d / L)
        rsq = (d ** 2).sum(axis=1)
        rsq[i] = 1e9
        got = set(int(j) for j in lst[head[i]:head[i] + nn[i]])
        # pairs within 1e-5 of the list cutoff may fall either way (fp32 displacement of the
        # wrapped image against float64 here); everything clearly inside must be there
        must = set(np.nonzero(rsq < (r_list * (1 - 1e-5)) ** 2)[0].tolist())
        may = set(np.nonzero(rsq < (r_list * (1 + 1e-5)) ** 2)[0].tolist())
        missing += len(must - got)
        extra += len(got - may)
    assert missing == 0 and extra == 0
