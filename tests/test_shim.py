"""The HOOMD-signature C++ shim (include/azp_hoomd_shim.h): compiles and links against the C ABI
without a GPU; on the GPU it reproduces a reference known-answer through
hoomd::md::kernel::gpu_compute_pair_forces<E>."""

import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = "/usr/local/cuda"


def build(tmp_path):
    import torch

    exe = str(tmp_path / "shim_driver")
    libdir = os.path.join(ROOT, "azplugins_b200")
    rt = os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib")
    rt = os.path.normpath(rt)
    cmd = ["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", CUDA + "/include",
           os.path.join(ROOT, "tests", "shim_driver.cc"), "-o", exe, "-L", libdir, "-lazp_b200",
           "-L", rt, "-l:libcudart.so.12", "-Wl,-rpath," + libdir, "-Wl,-rpath," + rt]
    subprocess.run(cmd, check=True)
    return exe


def test_shim_compiles_and_links(tmp_path):
    exe = build(tmp_path)
    out = subprocess.run([exe, "link"], capture_output=True, text=True, check=True).stdout
    assert out.strip() == "linked abi 1"


@pytest.mark.gpu
def test_shim_reproduces_reference_kats(tmp_path):
    """All three HOOMD driver templates, instantiated like the reference's *.cu.inc stubs do
    (src/PotentialPairGPUKernel.cu.inc:25-28, src/PotentialPairDPDThermoGPUKernel.cu.inc:21-24,
    src/AnisoPotentialPairGPUKernel.cu.inc:21-25), reproduce reference known answers
    (src/pytest/test_pair.py:76-85,177-186, src/pytest/test_pair_aniso.py:22-40)."""
    exe = build(tmp_path)
    lines = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()
    got = {ln.split()[0]: ln.split()[1:] for ln in lines}
    assert set(got) == {"hertz", "dpd", "morse"}
    for name, U, F in (("hertz", 0.0985, 0.5477), ("dpd", 0.25, 1.0)):
        out = got[name]
        assert out[0] == "rc" and out[1] == "0"
        f0x, e0, f1x, e1 = float(out[3]), float(out[7]), float(out[9]), float(out[11])
        assert abs(f0x + F) < 1.5e-4 and abs(f1x - F) < 1.5e-4, name
        assert abs(e0 - U / 2) < 1.5e-4 and abs(e1 - U / 2) < 1.5e-4, name
    out = got["morse"]
    assert out[1] == "0"
    f0 = [float(x) for x in out[3:6]]
    e0, f1x, e1 = float(out[7]), float(out[9]), float(out[11])
    t0 = [float(x) for x in out[13:16]]
    t1 = [float(x) for x in out[17:20]]
    for a, b in zip(f0, (11.75766, 2.46991, 3.70487)):
        assert abs(a - b) < 1.5e-4
    assert abs(f1x + 11.75766) < 1.5e-4
    assert abs(e0 + 0.41134 / 2) < 1.5e-4 and abs(e1 + 0.41134 / 2) < 1.5e-4
    for t in (t0, t1):
        for a, b in zip(t, (0.0, -0.08879, 0.05919)):
            assert abs(a - b) < 1.5e-4
