"""Generate tests/golden/ref_vectors.json from the REFERENCE evaluator headers.

Run in the build container (where /root/reference exists):

    make -C oracle ref && python tests/golden/make_golden.py

It drives oracle/_ref (the reference's own PairEvaluator*.h / DPDPairEvaluatorGeneralWeight.h /
AnisoPairEvaluatorTwoPatchMorse.h compiled in place) over a fixed, seeded set of single-pair
inputs and stores inputs + outputs (+ the packed param_type bytes of the reference
constructors). The fixture travels with the repository, so the restated oracle (oracle/port) and
the CUDA path can be checked against reference-generated numbers where /root/reference does not
exist. DPD random values come from the restated HOOMD RNG (the stream itself is unpinned by any
reference test; see oracle/hoomd_stub/hoomd/RandomNumbers.h).
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle  # noqa: E402

PARAM_SETS = {
    "PerturbedLennardJones": [
        dict(epsilon=1.0, sigma=1.0, attraction_scale_factor=0.5),
        dict(epsilon=2.0, sigma=0.85, attraction_scale_factor=0.25),
        dict(epsilon=0.7, sigma=1.3, attraction_scale_factor=0.0),
    ],
    "ExpandedYukawa": [
        dict(epsilon=1.0, kappa=1.0, delta=0.0),
        dict(epsilon=2.0, kappa=1.2, delta=0.15),
        dict(epsilon=3.0, kappa=1.5, delta=0.3),
    ],
    "Colloid": [
        dict(A=144.0, a_1=0.0, a_2=0.0, sigma=1.0),
        dict(A=144.0, a_1=0.0, a_2=5.0, sigma=1.0),
        dict(A=40.0, a_1=5.0, a_2=5.0, sigma=1.0),
        dict(A=100.0, a_1=1.5, a_2=0.75, sigma=1.05),
    ],
    "Hertz": [dict(epsilon=2.0), dict(epsilon=100.0)],
    "DPDGeneralWeight": [dict(A=25.0, gamma=4.5, s=2.0), dict(A=2.0, gamma=4.5, s=0.5)],
}
R_RANGE = {
    "PerturbedLennardJones": (0.8, 3.2, 3.0),
    "ExpandedYukawa": (0.7, 3.7, 3.5),
    "Hertz": (0.2, 1.6, 1.5),
    "DPDGeneralWeight": (0.3, 1.1, 1.0),
}
MORSE = [
    dict(M_d=1.8347, M_r=0.0302, r_eq=1.0043, omega=20.0, alpha=0.5, repulsion=True),
    dict(M_d=1.8341, M_r=0.0302, r_eq=1.0043, omega=5.0, alpha=0.40, repulsion=False),
]


def colloid_range(p):
    if p["a_1"] == 0 and p["a_2"] == 0:
        return 0.9, 3.2, 3.0
    if p["a_1"] != 0 and p["a_2"] != 0:
        s = p["a_1"] + p["a_2"]
        return s + 0.3, s + 0.7, s + 0.581
    a = max(p["a_1"], p["a_2"])
    return a + 0.6, a + 4.2, a + 4.0


def main():
    rng = np.random.default_rng(20260)
    out = {"_generator": "tests/golden/make_golden.py via oracle/_ref (reference headers in place)"}
    for bits, dt in ((32, np.float32), (64, np.float64)):
        ref = oracle.load("ref", dt)
        sec = {"pair": [], "dpd_thermo": [], "aniso": [], "params": [], "alpha": []}
        for name, plist in PARAM_SETS.items():
            for p in plist:
                lo, hi, rc = colloid_range(p) if name == "Colloid" else R_RANGE[name]
                sec["params"].append(dict(evaluator=name, params=p,
                                          bytes=ref.pack_params(name, p).tobytes().hex()))
                for shift in (False, True):
                    for r in rng.uniform(lo, hi, size=6):
                        r = float(dt(r))
                        ok, fdr, eng = ref.eval_pair(name, p, r * r, rc * rc, shift)
                        sec["pair"].append(dict(evaluator=name, params=p, rsq=r * r,
                                                rcutsq=rc * rc, shift=shift, evaluated=ok,
                                                force_divr=fdr, pair_eng=eng))
        for p in PARAM_SETS["DPDGeneralWeight"]:
            for _ in range(8):
                r = float(dt(rng.uniform(0.3, 1.05)))
                ti, tj = (int(x) for x in rng.integers(0, 2 ** 32 - 1, size=2))
                ts = int(rng.integers(0, 2 ** 40))
                seed = int(rng.integers(0, 65536))
                rdotv = float(dt(rng.normal()))
                ok, fdr, fc, eng = ref.eval_dpd_thermo(p, r * r, 1.0, seed, ti, tj, ts, 0.01,
                                                       rdotv, 1.0)
                sec["dpd_thermo"].append(dict(params=p, rsq=r * r, rcutsq=1.0, seed=seed,
                                              tag_i=ti, tag_j=tj, timestep=ts, dt=0.01,
                                              rdotv=rdotv, kT=1.0, evaluated=ok, force_divr=fdr,
                                              force_divr_cons=fc, pair_eng=eng))
                sec["alpha"].append(dict(seed=seed, tag_i=ti, tag_j=tj, timestep=ts,
                                         alpha=ref.dpd_alpha(seed, ti, tj, ts)))
        for p in MORSE:
            sec["params"].append(dict(evaluator="TwoPatchMorse", params=p,
                                      bytes=ref.pack_params("TwoPatchMorse", p).tobytes().hex()))
            for shift in (False, True):
                for _ in range(6):
                    d = rng.normal(size=3)
                    d *= rng.uniform(0.9, 1.7) / np.linalg.norm(d)
                    d = d.astype(dt).astype(np.float64)
                    qi = rng.normal(size=4)
                    qi = (qi / np.linalg.norm(qi)).astype(dt).astype(np.float64)
                    qj = rng.normal(size=4)
                    qj = (qj / np.linalg.norm(qj)).astype(dt).astype(np.float64)
                    ok, f, e, ti_, tj_ = ref.eval_aniso(p, d, qi, qj, 1.6 ** 2, shift)
                    sec["aniso"].append(dict(params=p, dr=d.tolist(), qi=qi.tolist(),
                                             qj=qj.tolist(), rcutsq=1.6 ** 2, shift=shift,
                                             evaluated=ok, force=f.tolist(), pair_eng=e,
                                             torque_i=ti_.tolist(), torque_j=tj_.tolist()))
        out["f%d" % bits] = sec
    with open(os.path.join(HERE, "ref_vectors.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote ref_vectors.json:", {k: {s: len(v) for s, v in sec.items()}
                                      for k, sec in out.items() if k != "_generator"})


if __name__ == "__main__":
    main()
