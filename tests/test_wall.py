"""Wall potentials wall.Colloid / wall.LJ93.

CPU: the oracle (restatement and the reference's evaluator headers compiled in place) against
the known answers of reference src/pytest/test_wall.py:20-58 and against each other on sphere /
cylinder / plane walls. GPU: the same known answers through the ``wall`` API and parity with the
oracle (LJ93 bit-exact: the kernel repeats the reference's IEEE operations; Colloid to the
rounding of ``log``)."""
import numpy as np
import pytest

from oracle import oracle

# (class name, params, position, energy, force_z)   reference src/pytest/test_wall.py:20-58
KATS = [
    ("Colloid", {"A": 1.0, "a": 1.0, "sigma": 1.0, "r_cut": 3.0}, (0, 0, 1.5), -0.0292, 0.8940),
    ("Colloid", {"A": 1.0, "a": 1.0, "sigma": 1.0, "r_cut": 3.0}, (0, 0, 3.5), 0.0, 0.0),
    ("LJ93", {"sigma": 1.0, "A": 1.0, "r_cut": 3.0}, (0, 0, 1.5), -0.2558, -0.5718),
    ("LJ93", {"sigma": 1.0, "A": 1.0, "r_cut": 3.0}, (0, 0, 3.5), 0.0, 0.0),
]


def _kinds():
    return [k for k in ("port", "ref") if oracle.available(k, 32)]


def _row(name, p, dtype):
    S = np.dtype(dtype).type
    rc = S(p["r_cut"])
    re = S(p.get("r_extrap", 0.0))
    if name == "Colloid":
        A, sigma = S(p["A"]), S(p["sigma"])
        s3 = sigma * sigma * sigma
        return [A * s3 * s3 / S(7560), A / S(6), S(p["a"]), rc * rc, re]
    sigma = S(p["sigma"])
    return [sigma * sigma * sigma, S(p["A"]), rc * rc, re]


@pytest.mark.parametrize("kind", _kinds())
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name,params,position,energy,fz", KATS)
def test_oracle_known_answers(kind, dtype, name, params, position, energy, fz):
    o = oracle.load(kind, dtype)
    pos = oracle.make_pos(np.array([position], dtype=np.float64), np.array([0]), dtype)
    r = o.wall_forces(name, pos, [_row(name, params, dtype)], planes=[[0, 0, 0, 0, 0, 1, 1]])
    np.testing.assert_array_almost_equal(r["force"][0, 3], energy, decimal=4)
    np.testing.assert_array_almost_equal(r["force"][0, :3], [0, 0, fz], decimal=4)


WALLSETS = {
    "plane": dict(planes=[[0, -6.0, 0, 0, 1, 0, 1], [0, 6.0, 0, 0, -1, 0, 0]]),
    "sphere": dict(spheres=[[7.0, 0.5, -0.25, 0.1, 1, 1], [2.0, 0, 0, 0, 0, 0]]),
    "cylinder": dict(cylinders=[[6.5, 0.2, 0.1, 0, 0, 0, 1, 1, 1], [1.5, 0, 0, 0, 0.6, 0.8, 0, 0, 1]]),
    "mixed": dict(planes=[[0, 0, -5.5, 0, 0, 1, 1]], spheres=[[7.5, 0, 0, 0, 1, 0]],
                  cylinders=[[6.0, 0, 0, 0, 1, 0, 0, 1, 1]]),
}


def _system(rng, N, dtype):
    xyz = rng.uniform(-7.0, 7.0, size=(N, 3)).astype(dtype).astype(np.float64)
    xyz[0] = 0.0  # on a sphere centre / cylinder axis
    return xyz, rng.integers(0, 2, N).astype(np.uint32)


PARAMS = {"Colloid": [{"A": 100.0, "a": 0.75, "sigma": 1.0, "r_cut": 3.0},
                      {"A": 40.0, "a": 1.25, "sigma": 0.9, "r_cut": 4.0}],
          "LJ93": [{"A": 2.0, "sigma": 1.0, "r_cut": 3.0}, {"A": 0.0, "sigma": 1.2, "r_cut": 2.5}]}
# the same with HOOMD's extrapolated mode switched on for the first type
PARAMS_EXTRAP = {"Colloid": [dict(PARAMS["Colloid"][0], r_extrap=1.1), PARAMS["Colloid"][1]],
                 "LJ93": [dict(PARAMS["LJ93"][0], r_extrap=0.9), PARAMS["LJ93"][1]]}


@pytest.mark.parametrize("kind", _kinds())
@pytest.mark.parametrize("name", ["Colloid", "LJ93"])
def test_oracle_extrapolated_mode_is_continuous_and_linear(kind, name):
    """r_extrap > 0: U and F continuous at r_extrap, F constant and U linear closer to (and
    behind) the wall."""
    o = oracle.load(kind, np.float64)
    p = PARAMS_EXTRAP[name][0]
    re = p["r_extrap"]
    z = np.array([re + 1e-9, re - 1e-9, 0.5 * re, 0.25 * re, 0.0, -0.3])
    pos = oracle.make_pos(np.c_[np.zeros_like(z), np.zeros_like(z), z], np.zeros(len(z), int), np.float64)
    r = o.wall_forces(name, pos, [_row(name, p, np.float64)], planes=[[0, 0, 0, 0, 0, 1, 1]])
    fz, u = r["force"][:, 2], r["force"][:, 3]
    assert abs(fz[0] - fz[1]) < 1e-6 * abs(fz[0]) and abs(u[0] - u[1]) < 1e-6 * max(1.0, abs(u[0]))
    assert np.allclose(fz[1:], fz[1], rtol=1e-12)  # constant force inside r_extrap and behind the wall
    # U(z) = U(r_e) + F (r_e - z): linear in z with slope -F
    assert np.allclose(u[2:], u[1] + fz[1] * (re - 1e-9 - z[2:]), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", ["Colloid", "LJ93"])
@pytest.mark.parametrize("wallset", sorted(WALLSETS))
@pytest.mark.parametrize("extrap", [False, True])
def test_port_equals_reference_headers(dtype, name, wallset, extrap):
    if not oracle.available("ref", 32):
        pytest.skip("oracle/_ref not built")
    xyz, typeid = _system(np.random.default_rng(3), 3000, dtype)
    pos = oracle.make_pos(xyz, typeid, dtype)
    rows = [_row(name, p, dtype) for p in (PARAMS_EXTRAP if extrap else PARAMS)[name]]
    a = oracle.load("port", dtype).wall_forces(name, pos, rows, **WALLSETS[wallset])
    b = oracle.load("ref", dtype).wall_forces(name, pos, rows, **WALLSETS[wallset])
    assert np.array_equal(a["force"], b["force"], equal_nan=True)
    assert np.array_equal(a["virial"], b["virial"], equal_nan=True)
    assert np.nanmax(np.abs(a["force"])) > 0  # (a colloid overlapping a wall has log(<0) = NaN energy, as in the reference)


# ------------------------------------------------------------------------------------------------
def _geometries(spec):
    from azplugins_b200 import wall

    out = [wall.Sphere(s[0], origin=s[1:4], inside=bool(s[4]), open=bool(s[5])) for s in spec.get("spheres", [])]
    out += [wall.Cylinder(c[0], axis=c[4:7], origin=c[1:4], inside=bool(c[7]), open=bool(c[8]))
            for c in spec.get("cylinders", [])]
    out += [wall.Plane(origin=p[0:3], normal=p[3:6], open=bool(p[6])) for p in spec.get("planes", [])]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name,params,position,energy,fz", KATS)
def test_gpu_known_answers(dtype, name, params, position, energy, fz):
    import azplugins_b200 as az

    state = az.State(az.Box.cube(20.0), ["A"], np.array([position], dtype=np.float64), dtype=dtype, device="cuda:0")
    pot = getattr(az.wall, name)(walls=[az.wall.Plane(origin=(0, 0, 0), normal=(0, 0, 1))])
    pot.params["A"] = params
    pot.attach(state).compute()
    for k, v in params.items():
        assert pot.params["A"][k] == v
    np.testing.assert_array_almost_equal(pot.energies, energy, decimal=4)
    np.testing.assert_array_almost_equal(pot.forces, [[0, 0, fz]], decimal=4)
    assert pot.cpp_class_name == "WallsPotential%sGPU" % name


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", ["Colloid", "LJ93"])
@pytest.mark.parametrize("wallset", sorted(WALLSETS))
@pytest.mark.parametrize("extrap", [False, True])
def test_gpu_parity_with_oracle(dtype, name, wallset, extrap):
    import azplugins_b200 as az

    xyz, typeid = _system(np.random.default_rng(11), 50021, dtype)
    state = az.State(az.Box.cube(20.0), ["A", "B"], xyz, typeid=typeid, dtype=dtype, device="cuda:0")
    walls = _geometries(WALLSETS[wallset])
    pot = getattr(az.wall, name)(walls=walls)
    for t, p in zip(["A", "B"], (PARAMS_EXTRAP if extrap else PARAMS)[name]):
        pot.params[t] = p
    pot.attach(state).compute()
    sph, cyl, pla = az.wall.walls_as_arrays(walls)
    rows = pot.param_table(["A", "B"], dtype)
    ref = oracle.load("best", dtype).wall_forces(name, state.pos.cpu().numpy(), rows, spheres=sph,
                                                 cylinders=cyl, planes=pla)
    got = np.concatenate([pot.forces, pot.energies[:, None]], axis=1)
    # a colloid overlapping a wall (r < a) has log(<0) = NaN energy in the reference as well:
    # those particles must be NaN here too, everything else is compared
    finite = np.isfinite(ref["force"]).all(axis=1)
    assert finite.mean() > 0.5
    assert np.array_equal(np.isnan(got), np.isnan(ref["force"]))
    if name == "LJ93":
        assert np.array_equal(got[finite], ref["force"][finite])
        assert np.array_equal(pot.virials[finite], ref["virial"].T[finite])
    else:
        tol = 2e-6 if dtype == np.float32 else 1e-13
        scale = np.abs(ref["force"][finite]).max(axis=0)
        assert (np.abs(got[finite] - ref["force"][finite]) <= tol * np.maximum(np.abs(ref["force"][finite]), scale * 1e-3)).all()
    assert np.abs(got[finite]).max() > 0


@pytest.mark.gpu
def test_gpu_wall_errors():
    import azplugins_b200 as az

    state = az.State(az.Box.cube(20.0), ["A"], np.zeros((4, 3)), dtype=np.float32, device="cuda:0")
    pot = az.wall.LJ93(walls=[az.wall.Plane((0, 0, -1), (0, 0, 1))])
    pot.attach(state)
    with pytest.raises(ValueError):
        pot.compute()  # params missing
    pot.params["A"] = dict(A=1.0, sigma=1.0, r_cut=3.0, r_extrap=-0.1)
    with pytest.raises(ValueError):
        pot.compute()
    with pytest.raises(ValueError):
        pot.params["A"] = dict(A=1.0, sigma=1.0)
    with pytest.raises(TypeError):
        az.wall.WallPotential(walls=[])
    bad = az.wall.LJ93(walls=["plane"])
    bad.params["A"] = dict(A=1.0, sigma=1.0, r_cut=3.0)
    with pytest.raises(TypeError):
        bad.attach(state).compute()
