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
def test_shim_reproduces_reference_kat(tmp_path):
    exe = build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    vals = dict(zip(out[0::2], out[1::2])) if False else out
    assert out[0] == "rc" and out[1] == "0"
    f0x, e0, f1x, e1 = float(out[3]), float(out[7]), float(out[9]), float(out[11])
    assert abs(f0x + 0.5477) < 1.5e-4 and abs(f1x - 0.5477) < 1.5e-4
    assert abs(e0 - 0.0985 / 2) < 1.5e-4 and abs(e1 - 0.0985 / 2) < 1.5e-4
