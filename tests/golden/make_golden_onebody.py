"""Generate tests/golden/ref_onebody.json from the REFERENCE headers of the one-body potentials.

Run in the build container (where /root/reference exists):

    make -C oracle ref && python tests/golden/make_golden_onebody.py

Drives oracle/_ref -- the reference's own PlanarBarrierEvaluator.h / SphericalBarrierEvaluator.h /
WallEvaluatorColloid.h / WallEvaluatorLJ93.h compiled in place (against oracle/hoomd_stub) -- over
fixed, seeded inputs and stores inputs + outputs, so that the restated oracle (port) and the CUDA
kernels can be checked against reference-generated numbers where /root/reference does not exist.
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle  # noqa: E402

L = [20.0, 24.0, 28.0]
BARRIER_PARAMS = [[50.0, 0.1], [200.0, -0.1], [0.0, 0.3]]
WALLS = dict(planes=[[0, 0, -5.5, 0, 0, 1, 1]], spheres=[[7.5, 0, 0, 0, 1, 0]],
             cylinders=[[6.0, 0, 0, 0, 1, 0, 0, 1, 1]])
WALL_PARAMS = {
    # rows {c_1, c_2, a, rcutsq, rextrap} / {sigma_3, A, rcutsq, rextrap} as wall.py stages them
    "Colloid": [dict(A=100.0, a=0.75, sigma=1.0, r_cut=3.0, r_extrap=0.0),
                dict(A=40.0, a=1.25, sigma=0.9, r_cut=4.0, r_extrap=1.6)],
    "LJ93": [dict(A=2.0, sigma=1.0, r_cut=3.0, r_extrap=0.0),
             dict(A=1.5, sigma=1.2, r_cut=2.5, r_extrap=0.9)],
}


def wall_row(name, p, dtype):
    S = np.dtype(dtype).type
    rc = S(p["r_cut"])
    if name == "Colloid":
        A, sigma = S(p["A"]), S(p["sigma"])
        s3 = sigma * sigma * sigma
        return [A * s3 * s3 / S(7560), A / S(6), S(p["a"]), rc * rc, S(p["r_extrap"])]
    sigma = S(p["sigma"])
    return [sigma * sigma * sigma, S(p["A"]), rc * rc, S(p["r_extrap"])]


def main():
    out = {"_generator": "tests/golden/make_golden_onebody.py (oracle/_ref = reference headers)",
           "L": L, "barrier_params": BARRIER_PARAMS, "walls": WALLS, "wall_params": WALL_PARAMS}
    rng = np.random.default_rng(20261)
    n = 96
    xyz = (rng.uniform(-0.5, 0.5, size=(n, 3)) * np.array(L))
    xyz[::7] *= 1.3  # some particles outside the box (wrapped by the barrier)
    typeid = rng.integers(0, 3, n)
    out["xyz"] = xyz.tolist()
    out["typeid"] = typeid.tolist()
    for dtype, key in ((np.float32, "f32"), (np.float64, "f64")):
        o = oracle.load("ref", dtype)
        pos = oracle.make_pos(xyz.astype(dtype).astype(np.float64), typeid, dtype)
        d = {}
        for geometry, loc in (("planar", 4.0), ("spherical", 6.5)):
            r = o.barrier_forces(geometry, loc, pos, BARRIER_PARAMS, L)
            d["barrier_" + geometry] = dict(location=loc, force=r["force"].astype(np.float64).tolist())
        wpos = oracle.make_pos(np.clip(xyz, -7.0, 7.0).astype(dtype).astype(np.float64), typeid % 2, dtype)
        for name in ("Colloid", "LJ93"):
            rows = [wall_row(name, p, dtype) for p in WALL_PARAMS[name]]
            r = o.wall_forces(name, wpos, rows, **WALLS)
            f = r["force"].astype(np.float64)
            d["wall_" + name] = dict(force=np.where(np.isfinite(f), f, 0.0).tolist(),
                                     finite=np.isfinite(f).all(axis=1).tolist(),
                                     virial=np.nan_to_num(r["virial"].astype(np.float64)).T.tolist())
        out[key] = d
    with open(os.path.join(HERE, "ref_onebody.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote ref_onebody.json")


if __name__ == "__main__":
    main()
