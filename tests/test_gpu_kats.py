"""GPU: the reference's own two-particle known-answer tests, run through the mirrored
``pair`` API and the C ABI on the CUDA path (reference src/pytest/test_pair.py:309-363 and
src/pytest/test_pair_aniso.py:113-168), fp32 and fp64."""

import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
pytestmark = pytest.mark.gpu


def assert_kat(actual, desired, dtype):
    actual = np.asarray(actual, dtype=np.float64)
    desired = np.asarray(desired, dtype=np.float64)
    rtol = 1e-5 if np.dtype(dtype) == np.float32 else 0.0
    assert (np.abs(actual - desired) <= 1.5e-4 + rtol * np.abs(desired)).all(), (actual, desired)


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("case", KATS["pair"], ids=lambda c: c["potential"])
def test_energy_and_force(case, dtype):
    import azplugins_b200 as az

    r_cut, r_buff = case["r_cut"], 0.4
    L = 2.1 * 2 * (r_cut + r_buff)
    d = case["distance"]
    state = az.State(az.Box.cube(L), ["A"], [[-d / 2, 0, 0], [d / 2, 0, 0]], dtype=dtype)
    cls = getattr(az.pair, case["potential"])
    extra = {}
    if cls is az.pair.DPDGeneralWeight:
        extra["kT"] = 0.0
    else:
        extra["mode"] = "shift" if case["shift"] else "none"
    potential = cls(nlist=az.nlist.Cell(buffer=r_buff), default_r_cut=r_cut, **extra)
    potential.params[("A", "A")] = case["params"]
    potential.attach(state)
    potential.compute()

    # parameters survive the trip through param_type on the device
    back = potential.get_params_from_device("A", "A")
    for k, v in case["params"].items():
        assert back[k] == pytest.approx(v, rel=1e-6 if dtype == np.float32 else 1e-14)
    e, f = case["energy"], case["force"]
    assert_kat(potential.energies, [0.5 * e, 0.5 * e], dtype)
    assert_kat(potential.forces, [[-f, 0, 0], [f, 0, 0]], dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("case", KATS["aniso"], ids=lambda c: "r_cut%g" % c["r_cut"])
def test_energy_force_and_torque(case, dtype):
    import azplugins_b200 as az

    state = az.State(az.Box.cube(20.0), ["A"], [[-0.5, -0.10, -0.15], [0.5, 0.10, 0.15]],
                     orientation=[[1, 0, 0, 0], [1, 0, 0, 0]], dtype=dtype)
    potential = az.pair.TwoPatchMorse(nlist=az.nlist.Cell(buffer=0.4), default_r_cut=case["r_cut"],
                                      mode="shift" if case["shift"] else "none")
    potential.params[("A", "A")] = case["params"]
    potential.attach(state)
    potential.compute()
    back = potential.get_params_from_device("A", "A")
    assert np.allclose([back[k] for k in case["params"]], [float(v) for v in case["params"].values()],
                       rtol=1e-6)
    e = case["energy"]
    assert_kat(potential.energies, [0.5 * e, 0.5 * e], dtype)
    if case["force"] is not None:
        F = np.array(case["force"])
        assert_kat(potential.forces, [-F, F], dtype)
    if case["torque"] is not None:
        T = np.array(case["torque"])
        assert_kat(potential.torques, [T, T], dtype)
