"""CPU: pin the oracle. (1) the reference's own known-answer vectors, (2) reference-generated
fixtures, (3) port == ref bit-for-bit where oracle/_ref is present, (4) Philox KATs."""

import json
import os

import numpy as np
import pytest

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
VEC = json.load(open(os.path.join(HERE, "golden", "ref_vectors.json")))

KINDS = ["port"] + (["ref"] if oracle.available("ref", 32) else [])
DTYPES = [np.float32, np.float64]


def assert_kat(actual, desired, dtype):
    """The reference asserts 4 decimals (|diff| < 1.5e-4) on a double build. The fp32 build adds
    its stated relative budget 1e-5 (r - delta = 0.05 in the Yukawa cases amplifies fp32 rounding
    of r to ~2e-6 relative on forces of ~4e2 .. 1.2e3)."""
    actual = np.asarray(actual, dtype=np.float64)
    desired = np.asarray(desired, dtype=np.float64)
    rtol = 1e-5 if np.dtype(dtype) == np.float32 else 0.0
    assert (np.abs(actual - desired) <= 1.5e-4 + rtol * np.abs(desired)).all(), (actual, desired)


def two_particle(orc, case, half):
    """The reference test's setup (src/pytest/test_pair.py:309-363): two particles on the x axis,
    box 2.1 * 2 (r_cut + 0.4), Cell(buffer=0.4); run through the restated HOOMD loop."""
    d = case["distance"]
    rc = case["r_cut"]
    L = 2.1 * 2 * (rc + 0.4)
    pos = oracle.make_pos([[-d / 2, 0, 0], [d / 2, 0, 0]], 0, orc.dtype)
    nn, nl, head = orc.build_nlist(pos, [L] * 3, rc + 0.4, half=half)
    name = case["potential"]
    table = orc.pack_table(name, 1, {(0, 0): case["params"]})
    mode = "shift" if case["shift"] else "none"
    if name == "DPDGeneralWeight":
        vel = np.zeros((2, 4), dtype=orc.dtype)
        return orc.dpd_forces(table, pos, vel, [0, 1], nn, nl, head, [L] * 3, rc, seed=1,
                              timestep=0, dt=0.001, kT=0.0, half=half)[0]
    return orc.pair_forces(name, table, pos, nn, nl, head, [L] * 3, rc, mode=mode, half=half)[0]


@pytest.mark.parametrize("half", [False, True], ids=["full", "half"])
@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "f64"])
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("case", KATS["pair"], ids=lambda c: c["potential"])
def test_reference_pair_kats(case, kind, dtype, half):
    orc = oracle.load(kind, dtype)
    f = two_particle(orc, case, half)
    e, fx = case["energy"], case["force"]
    assert_kat(f[:, 3], [0.5 * e, 0.5 * e], dtype)
    assert_kat(f[:, :3], [[-fx, 0, 0], [fx, 0, 0]], dtype)


@pytest.mark.parametrize("half", [False, True], ids=["full", "half"])
@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "f64"])
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("case", KATS["aniso"], ids=lambda c: "r_cut%g" % c["r_cut"])
def test_reference_aniso_kats(case, kind, dtype, half):
    orc = oracle.load(kind, dtype)
    L = 20.0  # HOOMD's two_particle_snapshot_factory default box
    pos = oracle.make_pos([[-0.5, -0.10, -0.15], [0.5, 0.10, 0.15]], 0, dtype)
    q = np.array([[1, 0, 0, 0], [1, 0, 0, 0]], dtype=dtype)
    rc = case["r_cut"]
    nn, nl, head = orc.build_nlist(pos, [L] * 3, rc + 0.4, half=half)
    table = orc.pack_table("TwoPatchMorse", 1, {(0, 0): case["params"]})
    f, t, _ = orc.aniso_forces(table, pos, q, nn, nl, head, [L] * 3, rc,
                               mode="shift" if case["shift"] else "none", half=half)
    e = case["energy"]
    assert_kat(f[:, 3], [0.5 * e, 0.5 * e], dtype)
    if case["force"] is not None:
        F = np.array(case["force"])
        assert_kat(f[:, :3], [-F, F], dtype)
    if case["torque"] is not None:
        T = np.array(case["torque"])
        assert_kat(t[:, :3], [T, T], dtype)


def test_philox_known_answers():
    """Philox4x32-10 KATs (Random123 kat_vectors)."""
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for kind in KINDS:
        orc = oracle.load(kind, np.float32)
        for ctr, key, want in kats:
            assert orc.philox(ctr, key).tolist() == want


@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "f64"])
def test_port_matches_reference_fixtures(dtype):
    """The restated evaluators reproduce reference-generated vectors bit-for-bit."""
    orc = oracle.load("port", dtype)
    sec = VEC["f%d" % (8 * np.dtype(dtype).itemsize)]
    for v in sec["pair"]:
        ok, fdr, eng = orc.eval_pair(v["evaluator"], v["params"], v["rsq"], v["rcutsq"], v["shift"])
        assert ok == v["evaluated"]
        assert fdr == v["force_divr"] and eng == v["pair_eng"], v
    for v in sec["dpd_thermo"]:
        got = orc.eval_dpd_thermo(v["params"], v["rsq"], v["rcutsq"], v["seed"], v["tag_i"],
                                  v["tag_j"], v["timestep"], v["dt"], v["rdotv"], v["kT"])
        assert got == (v["evaluated"], v["force_divr"], v["force_divr_cons"], v["pair_eng"])
    for v in sec["alpha"]:
        assert orc.dpd_alpha(v["seed"], v["tag_i"], v["tag_j"], v["timestep"]) == v["alpha"]
    for v in sec["aniso"]:
        ok, f, e, ti, tj = orc.eval_aniso(v["params"], v["dr"], v["qi"], v["qj"], v["rcutsq"], v["shift"])
        assert ok == v["evaluated"] and e == v["pair_eng"]
        assert f.tolist() == v["force"] and ti.tolist() == v["torque_i"] and tj.tolist() == v["torque_j"]
    for v in sec["params"]:
        assert orc.pack_params(v["evaluator"], v["params"]).tobytes().hex() == v["bytes"]


def random_system(n, rho, ntypes, dtype, seed):
    rng = np.random.default_rng(seed)
    L = (n / rho) ** (1 / 3)
    m = int(np.ceil(n ** (1 / 3)))
    g = (np.arange(m) + 0.5) * (L / m) - L / 2
    xyz = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)[:n]
    xyz = xyz + rng.uniform(-0.2, 0.2, size=xyz.shape) * (L / m)
    types = rng.integers(0, ntypes, size=n)
    return oracle.make_pos(xyz, types, dtype), L, rng


@pytest.mark.skipif(not oracle.available("ref", 32), reason="oracle/_ref not built")
@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "f64"])
def test_port_equals_reference_loops(dtype):
    """Whole-system loops: port and ref agree bit-for-bit (same driver, restated evaluators)."""
    port, ref = oracle.load("port", dtype), oracle.load("ref", dtype)
    pos, L, rng = random_system(1500, 0.6, 2, dtype, 7)
    nn, nl, head = port.build_nlist(pos, [L] * 3, 2.9, ntypes=2)
    yk = {(0, 0): dict(epsilon=1.0, kappa=1.0, delta=0.0), (0, 1): dict(epsilon=2.0, kappa=1.2, delta=0.15),
          (1, 1): dict(epsilon=3.0, kappa=1.5, delta=0.3)}
    plj = {k: dict(epsilon=1.0 + i, sigma=1.0 - 0.1 * i, attraction_scale_factor=0.5) for i, k in enumerate(yk)}
    hz = {k: dict(epsilon=5.0 + i) for i, k in enumerate(yk)}
    col = {(0, 0): dict(A=144.0, a_1=0, a_2=0, sigma=1.0), (0, 1): dict(A=144.0, a_1=0, a_2=0.4, sigma=0.3),
           (1, 1): dict(A=40.0, a_1=0.2, a_2=0.2, sigma=0.3)}
    for name, pp in (("ExpandedYukawa", yk), ("PerturbedLennardJones", plj), ("Hertz", hz), ("Colloid", col)):
        for mode in ("none", "shift", "xplor"):
            a = port.pair_forces(name, port.pack_table(name, 2, pp), pos, nn, nl, head, [L] * 3, 2.5,
                                 ntypes=2, r_on=2.0, mode=mode)
            b = ref.pair_forces(name, ref.pack_table(name, 2, pp), pos, nn, nl, head, [L] * 3, 2.5,
                                ntypes=2, r_on=2.0, mode=mode)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (name, mode)
    vel = np.zeros((len(pos), 4), dtype=dtype)
    vel[:, :3] = rng.normal(size=(len(pos), 3))
    tag = rng.permutation(len(pos)).astype(np.uint32)
    dp = {k: dict(A=25.0 + i, gamma=4.5, s=2.0 - 0.75 * i) for i, k in enumerate(yk)}
    a = port.dpd_forces(port.pack_table("DPDGeneralWeight", 2, dp), pos, vel, tag, nn, nl, head, [L] * 3,
                        1.9, 42, 123456789012, 0.01, 1.3, ntypes=2)
    b = ref.dpd_forces(ref.pack_table("DPDGeneralWeight", 2, dp), pos, vel, tag, nn, nl, head, [L] * 3,
                       1.9, 42, 123456789012, 0.01, 1.3, ntypes=2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    q = rng.normal(size=(len(pos), 4))
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(dtype)
    mp = {k: dict(M_d=1.8 + i, M_r=0.03 + 0.01 * i, r_eq=1.0, omega=20.0 - 5 * i, alpha=0.5, repulsion=bool(i % 2))
          for i, k in enumerate(yk)}
    for mode in ("none", "shift"):
        a = port.aniso_forces(port.pack_table("TwoPatchMorse", 2, mp), pos, q, nn, nl, head, [L] * 3, 1.8,
                              ntypes=2, mode=mode)
        b = ref.aniso_forces(ref.pack_table("TwoPatchMorse", 2, mp), pos, q, nn, nl, head, [L] * 3, 1.8,
                             ntypes=2, mode=mode)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "f64"])
def test_half_list_equals_full_list(dtype):
    """HOOMD's CPU classes use a half list + Newton's third law, the GPU classes a full list;
    the restated loop must give the same forces either way (up to summation order)."""
    orc = oracle.load("port", dtype)
    pos, L, _ = random_system(1200, 0.7, 1, dtype, 3)
    table = orc.pack_table("PerturbedLennardJones", 1,
                           {(0, 0): dict(epsilon=1.0, sigma=1.0, attraction_scale_factor=0.5)})
    full = orc.build_nlist(pos, [L] * 3, 2.9)
    half = orc.build_nlist(pos, [L] * 3, 2.9, half=True)
    assert full[0].sum() == 2 * half[0].sum()
    a = orc.pair_forces("PerturbedLennardJones", table, pos, *full, [L] * 3, 2.5)
    b = orc.pair_forces("PerturbedLennardJones", table, pos, *half, [L] * 3, 2.5, half=True)
    tol = 2e-4 if dtype == np.float32 else 1e-11
    assert np.abs(a[0] - b[0]).max() <= tol * max(1.0, np.abs(a[0]).max())
    assert np.abs(a[1] - b[1]).max() <= tol * max(1.0, np.abs(a[1]).max())


@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "f64"])
@pytest.mark.parametrize("nthreads", [1, 3, 8])
def test_domain_decomposed_half_list_equals_full_list(dtype, nthreads):
    """half="domains" (bench.py's CPU baseline: one half-list domain per thread, HOOMD's MPI
    decomposition restated over a full list) gives the full-list forces up to summation order,
    for all three loops and any number of domains."""
    orc = oracle.load("port", dtype)
    pos, L, rng = random_system(1500, 0.7, 2, dtype, 9)
    full = orc.build_nlist(pos, [L] * 3, 2.9, ntypes=2)
    tol = 3e-4 if dtype == np.float32 else 1e-11

    def close(a, b):
        return np.abs(a - b).max() <= tol * max(1.0, np.abs(a).max())

    yk = {(0, 0): dict(epsilon=1.0, kappa=1.0, delta=0.0), (0, 1): dict(epsilon=2.0, kappa=1.2, delta=0.15),
          (1, 1): dict(epsilon=3.0, kappa=1.5, delta=0.3)}
    t = orc.pack_table("ExpandedYukawa", 2, yk)
    a = orc.pair_forces("ExpandedYukawa", t, pos, *full, [L] * 3, 2.5, ntypes=2, mode="shift")
    b = orc.pair_forces("ExpandedYukawa", t, pos, *full, [L] * 3, 2.5, ntypes=2, mode="shift",
                        half="domains", nthreads=nthreads)
    assert close(a[0], b[0]) and close(a[1], b[1])
    vel = np.zeros((len(pos), 4), dtype=dtype)
    vel[:, :3] = rng.normal(size=(len(pos), 3))
    tag = rng.permutation(len(pos)).astype(np.uint32)
    dp = {k: dict(A=25.0, gamma=4.5, s=2.0) for k in yk}
    t = orc.pack_table("DPDGeneralWeight", 2, dp)
    a = orc.dpd_forces(t, pos, vel, tag, *full, [L] * 3, 1.9, 42, 1000, 0.01, 1.3, ntypes=2)
    b = orc.dpd_forces(t, pos, vel, tag, *full, [L] * 3, 1.9, 42, 1000, 0.01, 1.3, ntypes=2,
                       half="domains", nthreads=nthreads)
    assert close(a[0], b[0]) and close(a[1], b[1])
    q = rng.normal(size=(len(pos), 4))
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(dtype)
    mp = {k: dict(M_d=1.8, M_r=0.3, r_eq=1.0, omega=5.0, alpha=0.5, repulsion=True) for k in yk}
    t = orc.pack_table("TwoPatchMorse", 2, mp)
    a = orc.aniso_forces(t, pos, q, *full, [L] * 3, 1.8, ntypes=2)
    b = orc.aniso_forces(t, pos, q, *full, [L] * 3, 1.8, ntypes=2, half="domains", nthreads=nthreads)
    assert all(close(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "f64"])
def test_min_image_variants_agree(dtype):
    """Host (compare/subtract) and device (rint) minimum image agree away from exact ties."""
    orc = oracle.load("port", dtype)
    rng = np.random.default_rng(5)
    L = [10.0, 12.0, 9.0]
    for tilt in ((0, 0, 0), (0.2, -0.1, 0.3)):
        for _ in range(200):
            v = rng.uniform(-1, 1, 3) * np.array(L) * 0.95
            a = orc.min_image(v, L, tilt, rint=False)
            b = orc.min_image(v, L, tilt, rint=True)
            assert np.allclose(a, b, atol=1e-4 if dtype == np.float32 else 1e-12)
            if tilt == (0, 0, 0):
                assert (np.abs(a) <= np.array(L) / 2 + 1e-5).all()
