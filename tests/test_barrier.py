"""External harmonic barriers.

CPU: the oracle (own restatement and the reference's evaluator headers compiled in place) against
the known answers of reference src/pytest/test_external.py:84-230 and against each other.
GPU: the same known answers through the ``external`` API, and bit-exact parity with the oracle on
random systems (the kernel does the reference's IEEE operations one for one)."""
import numpy as np
import pytest

from oracle import oracle

kA, dB = 50.0, 2.0
kB = kA * dB ** 2
SPH_POS = np.array([[0, 0, 4.6], [0, 0, -5.4], [0, 5.6, 0], [6.6, 0, 0]], dtype=np.float64)
PLA_POS = np.array([[1, 4.6, 1], [-1, 5.4, 1], [1, 5.6, -1], [-1, 6.6, -1]], dtype=np.float64)
TYPEID = np.array([0, 1, 0, 0], dtype=np.uint32)
# (geometry, location, kB) -> expected forces, energies  (test_external.py:113-149, 179-229)
KATS = [
    ("spherical", 5.0, kB, [[0, 0, 0], [0, 0, kB * 0.5], [0, -kA * 0.5, 0], [-kA * 1.5, 0, 0]],
     [0, 0.5 * kB * 0.25, 0.5 * kA * 0.25, 0.5 * kA * 2.25]),
    ("spherical", 4.0, 0.0, [[0, 0, -kA * 0.5], [0, 0, 0], [0, -kA * 1.5, 0], [-kA * 2.5, 0, 0]],
     [0.5 * kA * 0.25, 0, 0.5 * kA * 2.25, 0.5 * kA * 6.25]),
    ("planar", 5.0, kB, [[0, 0, 0], [0, -kB * 0.5, 0], [0, -kA * 0.5, 0], [0, -kA * 1.5, 0]],
     [0, 0.5 * kB * 0.25, 0.5 * kA * 0.25, 0.5 * kA * 2.25]),
    ("planar", 4.0, 0.0, [[0, -kA * 0.5, 0], [0, 0, 0], [0, -kA * 1.5, 0], [0, -kA * 2.5, 0]],
     [0.5 * kA * 0.25, 0, 0.5 * kA * 2.25, 0.5 * kA * 6.25]),
]


def _kinds():
    return [k for k in ("port", "ref") if oracle.available(k, 32)]


def _random_system(rng, N, L, dtype, outside=True):
    xyz = rng.uniform(-0.5, 0.5, size=(N, 3)) * np.asarray(L)
    if outside:  # some particles drifted out of the box by less than one image
        sel = rng.random(N) < 0.2
        xyz[sel] += rng.uniform(-0.3, 0.3, size=(int(sel.sum()), 3)) * np.asarray(L)
    typeid = rng.integers(0, 3, N).astype(np.uint32)
    return xyz.astype(dtype).astype(np.float64), typeid


@pytest.mark.parametrize("kind", _kinds())
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("geometry,location,kb,forces,energies", KATS)
def test_oracle_known_answers(kind, dtype, geometry, location, kb, forces, energies):
    o = oracle.load(kind, dtype)
    xyz = SPH_POS if geometry == "spherical" else PLA_POS
    r = o.barrier_forces(geometry, location, oracle.make_pos(xyz, TYPEID, dtype),
                         [[kA, 0.1], [kb, -0.1]], [20, 20, 20])
    np.testing.assert_allclose(r["force"][:, :3], forces, atol=1e-4)
    np.testing.assert_allclose(r["force"][:, 3], energies, atol=1e-4)
    assert not r["virial"].any()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("geometry", ["planar", "spherical"])
def test_port_equals_reference_headers(dtype, geometry):
    if not oracle.available("ref", 32):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    for L, tilt in (([20, 24, 28], (0, 0, 0)), ([20, 24, 28], (0.1, -0.2, 0.15))):
        xyz, typeid = _random_system(rng, 4000, L, dtype)
        pos = oracle.make_pos(xyz, typeid, dtype)
        par = [[50.0, 0.1], [200.0, -0.1], [0.0, 0.3]]
        a = oracle.load("port", dtype).barrier_forces(geometry, 6.5, pos, par, L, tilt)
        b = oracle.load("ref", dtype).barrier_forces(geometry, 6.5, pos, par, L, tilt)
        assert np.array_equal(a["force"], b["force"])


@pytest.mark.parametrize("kind", _kinds())
def test_oracle_rejects_invalid_location(kind):
    o = oracle.load(kind, np.float32)
    pos = oracle.make_pos(PLA_POS, TYPEID, np.float32)
    with pytest.raises(RuntimeError, match="Barrier position is invalid"):
        o.barrier_forces("planar", 10.0, pos, [[1, 0], [1, 0]], [20, 20, 20])  # H == hi.y
    with pytest.raises(RuntimeError, match="Barrier position is invalid"):
        o.barrier_forces("spherical", 10.5, pos, [[1, 0], [1, 0]], [20, 20, 20])  # 2R > L
    o.barrier_forces("planar", -10.0, pos, [[1, 0], [1, 0]], [20, 20, 20])  # H == lo.y is inside
    o.barrier_forces("spherical", 10.0, pos, [[1, 0], [1, 0]], [20, 20, 20])


# ------------------------------------------------------------------------------------------------
def _gpu_barrier(geometry, location, xyz, typeid, types, params, L, dtype, tilt=(0, 0, 0)):
    import azplugins_b200 as az

    box = az.Box(L[0], L[1], L[2], xy=tilt[0], xz=tilt[1], yz=tilt[2])
    state = az.State(box, types, xyz, typeid=typeid, dtype=dtype, device="cuda:0")
    cls = az.external.PlanarHarmonicBarrier if geometry == "planar" else az.external.SphericalHarmonicBarrier
    b = cls(location=location)
    for t, p in zip(types, params):
        b.params[t] = dict(k=p[0], offset=p[1])
    return b.attach(state), state


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("geometry,location,kb,forces,energies", KATS)
def test_gpu_known_answers(dtype, geometry, location, kb, forces, energies):
    xyz = SPH_POS if geometry == "spherical" else PLA_POS
    # the reference test moves the barrier with a custom variant: 5.0 up to step 1, then 4.0
    variant = lambda timestep: 5.0 if timestep <= 1 else 4.0  # noqa: E731
    b, _ = _gpu_barrier(geometry, variant, xyz, TYPEID, ["A", "B"], [[kA, 0.1], [kb, -0.1]],
                        [20, 20, 20], dtype)
    b.compute(timestep=1 if location == 5.0 else 3)
    np.testing.assert_allclose(b.forces, forces, atol=1e-4)
    np.testing.assert_allclose(b.energies, energies, atol=1e-4)
    assert not b.virials.any()
    assert b.cpp_class_name in ("PlanarHarmonicBarrierGPU", "SphericalHarmonicBarrierGPU")


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("geometry", ["planar", "spherical"])
@pytest.mark.parametrize("N", [1, 3, 4, 1025, 200003])
def test_gpu_bit_exact_against_oracle(dtype, geometry, N):
    rng = np.random.default_rng(100 + N)
    o = oracle.load("best", dtype)
    for L, tilt in (([20, 24, 28], (0, 0, 0)), ([20, 24, 28], (0.1, -0.2, 0.15))):
        xyz, typeid = _random_system(rng, N, L, dtype)
        par = [[50.0, 0.1], [200.0, -0.1], [0.0, 0.3]]
        b, state = _gpu_barrier(geometry, 6.5, xyz, typeid, ["A", "B", "C"], par, L, dtype, tilt)
        b._virial.fill_(3.0)
        b.compute()
        ref = o.barrier_forces(geometry, 6.5, state.pos.cpu().numpy(), par, L, tilt)
        got = np.concatenate([b.forces, b.energies[:, None]], axis=1)
        assert np.array_equal(got, ref["force"]), np.abs(got - ref["force"]).max()
        assert not b.virials.any()


@pytest.mark.gpu
def test_gpu_invalid_location_and_errors():
    import azplugins_b200 as az

    b, state = _gpu_barrier("planar", 10.0, PLA_POS, TYPEID, ["A", "B"], [[1, 0], [1, 0]], [20, 20, 20], np.float32)
    with pytest.raises(RuntimeError, match="Barrier position is invalid"):
        b.compute()
    b.location = 3.0
    del b.params["B"]
    with pytest.raises(ValueError):
        b.compute()
    with pytest.raises(TypeError):
        az.external.HarmonicBarrier(location=1.0)
    with pytest.raises(ValueError):
        b.params["B"] = dict(k=1.0)
