"""One-body potentials against the fixture generated from the reference's own headers
(tests/golden/ref_onebody.json, made by tests/golden/make_golden_onebody.py in the build container):
the restated oracle on the CPU, the CUDA kernels on the GPU."""
import json
import os

import numpy as np
import pytest

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "ref_onebody.json")))
import sys  # noqa: E402

sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_onebody import wall_row  # noqa: E402

DT = {"f32": np.float32, "f64": np.float64}


def _inputs(key):
    dtype = DT[key]
    xyz = np.array(G["xyz"])
    typeid = np.array(G["typeid"])
    return dtype, xyz.astype(dtype).astype(np.float64), typeid, np.clip(xyz, -7.0, 7.0).astype(dtype).astype(np.float64)


@pytest.mark.parametrize("key", ["f32", "f64"])
def test_port_oracle_reproduces_reference_fixture(key):
    dtype, xyz, typeid, wxyz = _inputs(key)
    o = oracle.load("port", dtype)
    pos = oracle.make_pos(xyz, typeid, dtype)
    for geometry in ("planar", "spherical"):
        ref = G[key]["barrier_" + geometry]
        r = o.barrier_forces(geometry, ref["location"], pos, G["barrier_params"], G["L"])
        assert np.array_equal(r["force"].astype(np.float64), np.array(ref["force"]))
    wpos = oracle.make_pos(wxyz, typeid % 2, dtype)
    for name in ("Colloid", "LJ93"):
        ref = G[key]["wall_" + name]
        rows = [wall_row(name, p, dtype) for p in G["wall_params"][name]]
        r = o.wall_forces(name, wpos, rows, **G["walls"])
        fin = np.array(ref["finite"])
        assert np.array_equal(np.isfinite(r["force"]).all(axis=1), fin)
        assert np.array_equal(r["force"][fin].astype(np.float64), np.array(ref["force"])[fin])
        assert np.array_equal(r["virial"].T[fin].astype(np.float64), np.array(ref["virial"])[fin])


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["f32", "f64"])
def test_gpu_reproduces_reference_fixture(key):
    import azplugins_b200 as az

    dtype, xyz, typeid, wxyz = _inputs(key)
    L = G["L"]
    state = az.State(az.Box(*L), ["A", "B", "C"], xyz, typeid=typeid, dtype=dtype, device="cuda:0")
    for geometry, cls in (("planar", az.external.PlanarHarmonicBarrier), ("spherical", az.external.SphericalHarmonicBarrier)):
        ref = G[key]["barrier_" + geometry]
        b = cls(location=ref["location"])
        for t, (k, off) in zip(["A", "B", "C"], G["barrier_params"]):
            b.params[t] = dict(k=k, offset=off)
        b.attach(state).compute()
        got = np.c_[b.forces, b.energies].astype(np.float64)
        assert np.array_equal(got, np.array(ref["force"]))  # bit-identical to the reference headers
    wstate = az.State(az.Box(*L), ["A", "B"], wxyz, typeid=typeid % 2, dtype=dtype, device="cuda:0")
    w = G["walls"]
    walls = [az.wall.Sphere(s[0], origin=s[1:4], inside=bool(s[4]), open=bool(s[5])) for s in w["spheres"]]
    walls += [az.wall.Cylinder(c[0], axis=c[4:7], origin=c[1:4], inside=bool(c[7]), open=bool(c[8])) for c in w["cylinders"]]
    walls += [az.wall.Plane(origin=p[0:3], normal=p[3:6], open=bool(p[6])) for p in w["planes"]]
    for name in ("Colloid", "LJ93"):
        ref = G[key]["wall_" + name]
        pot = getattr(az.wall, name)(walls=walls)
        for t, p in zip(["A", "B"], G["wall_params"][name]):
            pot.params[t] = p
        pot.attach(wstate).compute()
        fin = np.array(ref["finite"])
        got = np.c_[pot.forces, pot.energies].astype(np.float64)[fin]
        want = np.array(ref["force"])[fin]
        if name == "LJ93":
            assert np.array_equal(got, want)
        else:
            tol = 2e-6 if key == "f32" else 1e-13
            assert (np.abs(got - want) <= tol * np.maximum(np.abs(want), 1e-3 * np.abs(want).max(axis=0))).all()
