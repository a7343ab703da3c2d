"""CPU: the C-ABI library loads and exports every symbol include/azp_b200.h declares (no compute
calls), host-side parameter packing reproduces the reference constructors, and the Python mirror
of hoomd.azplugins.pair keeps the reference's API surface."""

import ctypes
import json
import os
import re

import numpy as np
import pytest

from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VEC = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_vectors.json")))


def test_library_exports_every_declared_symbol():
    from azplugins_b200 import _lib

    header = open(os.path.join(ROOT, "include", "azp_b200.h")).read()
    declared = set(re.findall(r"\b(azp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert getattr(lib, name) is not None
    assert _lib.lib.azp_abi_version() == 1


def test_struct_layout_matches_header_sizes():
    """ctypes mirrors of the C structs have the sizes the compiler gives them."""
    import subprocess
    import tempfile

    from azplugins_b200 import _lib

    src = ('#include "azp_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", '
           'sizeof(azp_box), sizeof(azp_pair_args), sizeof(azp_nlist_args), sizeof(azp_barrier_args), '
           'sizeof(azp_wall_args), sizeof(azp_md_args));return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o",
                        os.path.join(d, "s")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout
    sizes = [int(x) for x in out.split()]
    assert sizes == [ctypes.sizeof(_lib.AzpBox), ctypes.sizeof(_lib.AzpPairArgs),
                     ctypes.sizeof(_lib.AzpNlistArgs), ctypes.sizeof(_lib.AzpBarrierArgs),
                     ctypes.sizeof(_lib.AzpWallArgs), ctypes.sizeof(_lib.AzpMdArgs)]


def test_evaluator_names_and_param_sizes():
    """getName() strings and sizeof(param_type) of the reference."""
    from azplugins_b200 import _lib, kernels

    names = [_lib.lib.azp_evaluator_name(e).decode() for e in range(6)]
    assert names == ["PerturbedLennardJones", "ExpandedYukawa", "colloid", "hertz", "dpd_gen",
                     "TwoPatchMorse"]
    assert [(kernels.param_size(e, 32), kernels.param_size(e, 64)) for e in range(6)] == \
        [(16, 32), (16, 32), (16, 32), (4, 8), (16, 32), (24, 48)]
    if oracle.available("ref", 32):
        ref = oracle.load("ref", np.float32)
        for name, ev in oracle.EVALUATORS.items():
            assert ref.name(name) == names[ev]
            assert ref.param_size(name) == kernels.param_size(ev, 32)


@pytest.mark.parametrize("bits", [32, 64])
def test_param_pack_matches_reference_constructor_bytes(bits):
    """azp_param_pack == the reference's param_type(pybind11::dict) bytes (fixture generated from
    the reference headers), and unpack round-trips like asDict()/toPython()."""
    from azplugins_b200 import kernels

    for v in VEC["f%d" % bits]["params"]:
        ev = oracle.EVALUATORS[v["evaluator"]]
        fields = [float(v["params"][k]) for k in oracle.PARAM_FIELDS[v["evaluator"]]]
        raw = kernels.pack_params(ev, bits, fields)
        assert raw.tobytes().hex() == v["bytes"], v
        back = kernels.unpack_params(ev, bits, raw)
        assert np.allclose(back, fields, rtol=1e-6 if bits == 32 else 1e-14)
    if oracle.available("ref", bits):
        ref = oracle.load("ref", np.float32 if bits == 32 else np.float64)
        p = dict(epsilon=1.3, sigma=0.97, attraction_scale_factor=0.25)
        raw = kernels.pack_params(0, bits, [p["epsilon"], p["sigma"], p["attraction_scale_factor"]])
        assert np.array_equal(raw, ref.pack_params("PerturbedLennardJones", p))
        d = ref.unpack_params("PerturbedLennardJones", raw)
        back = kernels.unpack_params(0, bits, raw)
        assert back[0] == d["epsilon"] and back[1] == d["sigma"]


def test_dpd_rng_host_matches_oracle():
    from azplugins_b200 import kernels

    assert kernels.philox4x32_10([0, 0, 0, 0], [0, 0]).tolist() == \
        [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    for bits, dt in ((32, np.float32), (64, np.float64)):
        orc = oracle.load("best", dt)
        for v in VEC["f%d" % bits]["alpha"]:
            a = kernels.dpd_alpha(bits, v["seed"], v["tag_i"], v["tag_j"], v["timestep"])
            assert a == v["alpha"] == orc.dpd_alpha(v["seed"], v["tag_i"], v["tag_j"], v["timestep"])
            assert a == kernels.dpd_alpha(bits, v["seed"], v["tag_j"], v["tag_i"], v["timestep"])


def test_pair_api_surface():
    """Class names, C++ class-name mapping, accepted modes, ctor signatures, param schemas
    (reference src/pair.py:107-118, 214-239, 288-297, 341-351, 414-426, 508-525)."""
    from azplugins_b200 import nlist, pair

    nl = nlist.Cell(buffer=0.4)
    expect = {
        "Colloid": ("PotentialPairColloid", ("none", "shift", "xplor"), ["A", "a_1", "a_2", "sigma"]),
        "ExpandedYukawa": ("PotentialPairExpandedYukawa", ("none", "shift", "xplor"), ["epsilon", "kappa", "delta"]),
        "Hertz": ("PotentialPairHertz", ("none", "shift", "xplor"), ["epsilon"]),
        "PerturbedLennardJones": ("PotentialPairPerturbedLennardJones", ("none", "shift", "xplor"),
                                  ["epsilon", "sigma", "attraction_scale_factor"]),
    }
    for name, (cpp, modes, keys) in expect.items():
        cls = getattr(pair, name)
        pot = cls(nl, default_r_cut=3.0, default_r_on=0, mode="none")
        assert pot._cpp_class_name == cpp and pot.cpp_class_name == cpp + "GPU"
        assert pot._accepted_modes == modes and list(pot._param_schema) == keys
        with pytest.raises(KeyError):
            pot.params[("A", "A")] = {"bogus": 1.0}
        pot.params[("A", "B")] = {k: 1.0 for k in keys}
        assert pot.params[("B", "A")] == {k: 1.0 for k in keys}
        assert pot.r_cut[("A", "B")] == 3.0 and pot.r_on[("A", "A")] == 0.0
        pot.r_cut[("B", "A")] = 2.0
        assert pot.r_cut[("A", "B")] == 2.0
    dpd = pair.DPDGeneralWeight(nlist=nl, kT=1.5, default_r_cut=1.0)
    assert dpd._cpp_class_name == "PotentialPairDPDThermoGeneralWeight" and dpd._accepted_modes == ("none",)
    assert dpd.cpp_class_name == "PotentialPairDPDThermoGeneralWeightGPU" and dpd._kT(0) == 1.5
    assert pair.DPDGeneralWeight(nl, kT=lambda t: 2.0 + t)._kT(3) == 5.0
    m2p = pair.TwoPatchMorse(nlist=nl, default_r_cut=1.6, mode="shift")
    assert m2p._cpp_class_name == "AnisoPotentialPairTwoPatchMorse" and m2p._accepted_modes == ("none", "shift")
    m2p.params[("A", "A")] = dict(M_d=1.8, M_r=0.03, r_eq=1.0, omega=20, alpha=0.5, repulsion=True)
    assert m2p.params[("A", "A")]["repulsion"] is True
    with pytest.raises(ValueError):
        pair.TwoPatchMorse(nlist=nl, mode="xplor")
    with pytest.raises(TypeError):
        pair.Hertz(nlist="not a list")


def test_no_cpu_fallback():
    """Attaching to a non-CUDA state must fail loudly: there is no CPU implementation."""
    import azplugins_b200 as az
    from azplugins_b200 import _lib

    state = az.State(az.Box.cube(10.0), ["A"], [[0, 0, 0], [1, 0, 0]], device="cpu")
    nl = az.nlist.Cell(buffer=0.4)
    pot = az.pair.Hertz(nlist=nl, default_r_cut=1.5)
    pot.params[("A", "A")] = dict(epsilon=1.0)
    with pytest.raises(_lib.AzpError):
        pot.attach(state)
    with pytest.raises(_lib.AzpError):
        nl.build(state)
    # the product package never imports the oracle
    import sys

    # the one-body potentials and the integrator refuse a CPU state as well
    bar = az.external.PlanarHarmonicBarrier(location=1.0)
    wall = az.wall.LJ93(walls=[az.wall.Plane((0, 0, 0), (0, 0, 1))])
    for obj in (bar, wall, az.md.Integrator(dt=0.001, forces=[bar])):
        with pytest.raises(_lib.AzpError):
            obj.attach(state)
    for mod in ("azplugins_b200.pair", "azplugins_b200.kernels", "azplugins_b200.nlist",
                "azplugins_b200.slices", "azplugins_b200.synth", "azplugins_b200.external",
                "azplugins_b200.wall", "azplugins_b200.md"):
        __import__(mod)
        src = open(sys.modules[mod].__file__).read()
        assert "import oracle" not in src and "from oracle" not in src


def test_synthetic_workloads():
    from azplugins_b200 import synth

    wl = synth.config2(N=8000)
    assert wl.N == 8000 and set(np.unique(wl.typeid)) == {0, 1}
    L = wl.box.Lx
    assert abs(wl.N / L ** 3 - 0.5) < 1e-9 and np.abs(wl.position).max() <= L / 2
    assert sorted(wl.tag.tolist()) == list(range(8000))
    assert abs(synth.config2().n_bar - 124.2) < 0.1 and abs(synth.config1().n_bar - 131.7) < 0.1
    wl5 = synth.config5(N=2000)
    assert np.allclose(np.linalg.norm(wl5.orientation, axis=1), 1.0)
    wl3 = synth.config3(N=120000)
    assert (wl3.typeid == 1).sum() == 60 and abs(wl3.N - 120000) / 120000 < 0.05
    # Morton order keeps consecutive particles close
    d = np.linalg.norm(np.diff(wl.position, axis=0), axis=1)
    assert np.median(d) < 2.5


def test_external_and_wall_host_logic():
    """API surface of the external / wall mirrors (reference src/external.py, src/wall.py) and the
    host-side staging: wall-list layout, parameter rows with the reference constructors'
    roundings, the barrier validity rule."""
    import azplugins_b200 as az
    from azplugins_b200 import _lib, wall

    # class-name mapping and parameter schema
    bar = az.external.SphericalHarmonicBarrier(location=3.0)
    assert bar.cpp_class_name == "SphericalHarmonicBarrierGPU" and bar.location(7) == 3.0
    bar.params["A"] = dict(k=10.0, offset=0.5)
    assert bar.params["A"] == dict(k=10.0, offset=0.5)
    with pytest.raises(ValueError):
        bar.params["A"] = dict(k=1.0)
    ramp = az.external.PlanarHarmonicBarrier(location=lambda t: 5.0 - 0.001 * t)
    assert ramp.location(1000) == 4.0
    col = wall.Colloid(walls=[])
    assert col.cpp_class_name == "WallsPotentialColloidGPU"
    col.params["A"] = dict(A=1.0, a=1.0, sigma=1.0, r_cut=3.0)
    assert col.params["A"]["r_extrap"] == 0.0
    # parameter rows: c_1 = A sigma^6 / 7560, c_2 = A / 6 in Scalar arithmetic (WallEvaluatorColloid.h:36-45)
    for dtype in (np.float32, np.float64):
        S = np.dtype(dtype).type
        row = col.param_table(["A"], dtype)[0]
        assert row.dtype == dtype and row.shape == (5,)
        assert row[0] == S(1.0) / S(7560) and row[1] == S(1.0) / S(6) and row[3] == S(9.0)
        assert col.param_table(["A"], dtype).nbytes == _lib.lib.azp_wall_param_size(_lib.WALL_COLLOID, 8 * row.itemsize)
        lj = wall.LJ93(walls=[])
        lj.params["A"] = dict(A=2.0, sigma=1.5, r_cut=2.0, r_extrap=0.5)
        r2 = lj.param_table(["A"], dtype)[0]
        assert r2[0] == S(1.5) * S(1.5) * S(1.5) and r2[1] == S(2.0) and r2[2] == S(4.0) and r2[3] == S(0.5)
        assert lj.param_table(["A"], dtype).nbytes == _lib.lib.azp_wall_param_size(_lib.WALL_LJ93, 8 * row.itemsize)
        # wall list layout = the library's struct
        walls = [wall.Sphere(2.0, origin=(1, 2, 3), inside=False), wall.Cylinder(1.5, axis=(0, 0, 2)),
                 wall.Plane((0, 0, -1), (0, 3, 4), open=False)]
        blob = wall.pack_walls(walls, dtype)
        assert len(blob) == _lib.lib.azp_walls_size(8 * row.itemsize)
        head = np.frombuffer(blob[:16], dtype=np.uint32)
        assert list(head[:3]) == [1, 1, 1]
        sph, cyl, pla = wall.walls_as_arrays(walls)
        assert sph == [[2.0, 1.0, 2.0, 3.0, False, True]] and cyl[0][4:7] == [0.0, 0.0, 1.0]
        assert np.allclose(pla[0][3:6], [0.0, 0.6, 0.8]) and pla[0][6] is False
    with pytest.raises(ValueError):
        wall.pack_walls([wall.Plane((0, 0, 0), (0, 0, 1))] * 61, np.float32)
    with pytest.raises(ValueError):
        wall.Plane((0, 0, 0), (0, 0, 0))
    # barrier validity (PlanarBarrierEvaluator.h:54-59, SphericalBarrierEvaluator.h:56-62)
    box = az.Box.cube(20.0).to_c()
    valid = _lib.lib.azp_harmonic_barrier_valid
    for bits in (32, 64):
        assert valid(_lib.BARRIER_PLANAR, bits, -10.0, ctypes.byref(box)) == 1
        assert valid(_lib.BARRIER_PLANAR, bits, 10.0, ctypes.byref(box)) == 0
        assert valid(_lib.BARRIER_SPHERICAL, bits, 10.0, ctypes.byref(box)) == 1
        assert valid(_lib.BARRIER_SPHERICAL, bits, 10.01, ctypes.byref(box)) == 0
        assert valid(_lib.BARRIER_SPHERICAL, bits, -0.1, ctypes.byref(box)) == 0
    # integrator argument checks
    with pytest.raises(ValueError):
        az.md.Integrator(dt=0.001, forces=[bar], methods=[object()])


def test_config3_hole_removal_equals_the_brute_force_definition():
    """synth.config3 removes the solvent inside r < 5.9 of every colloid through a cell binning;
    the result must be the one of testing every solvent particle against every colloid."""
    from azplugins_b200 import synth

    wl = synth.config3(N=60000, n_colloid=3000)  # 45 colloids
    L = wl.box.L[0]
    col = wl.position[wl.typeid == 1]
    sol = wl.position[wl.typeid == 0]
    assert len(col) == 45 and len(sol) > 50000
    # no solvent particle is inside a hole ...
    for c in col:
        d = sol - c
        d -= L * np.round(d / L)
        assert ((d ** 2).sum(axis=1) >= 5.9 ** 2).all()
    # ... and the holes are not larger than that: the solvent density outside them is the
    # lattice's (every removed particle was inside some hole)
    n_lat = int(round(0.7 * L ** 3))
    removed = n_lat - len(sol)
    expect = 45 * 4.0 / 3.0 * np.pi * 5.9 ** 3 * 0.7
    assert abs(removed - expect) < 0.03 * expect
