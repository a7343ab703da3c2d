"""Host logic of the measurement chain (no GPU): the ncu-derived constants bench.py reports are
tied to the kernel sources they were captured with, and the registration tool produces entries
bench.py accepts."""

import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_source_hash_is_stable_and_sensitive(tmp_path):
    import srchash

    h = srchash.kernel_source_hash()
    assert h == srchash.kernel_source_hash() and len(h) == 16
    # the hash covers the files that determine the pair-force SASS
    base = os.path.join(ROOT, "azplugins_b200", "csrc")
    for name in ("pair_kernels.cuh", "azp_core.cuh", "launch.cuh", "capi.cu", "Makefile", "inst_yukawa.cu"):
        assert os.path.exists(os.path.join(base, name))


def test_constants_are_reported_only_for_the_kernels_they_were_captured_with(monkeypatch):
    import bench
    import srchash

    here = srchash.kernel_source_hash()
    table = json.load(open(os.path.join(ROOT, "profiles", "ncu_constants.json")))
    # no SASS table for the loaded library: the source hash decides
    monkeypatch.setattr(srchash, "kernel_sass_hashes", lambda lib_path=None: {})
    entry, why, h = bench.ncu_constants("C2")
    assert h == here
    if table["C2"]["kernel_sources"] == here:
        assert entry is not None and entry["dram_bytes"] > 5e8 and os.path.exists(os.path.join(ROOT, why))
    else:
        assert entry is None and "stale" in why
    monkeypatch.setattr(srchash, "kernel_source_hash", lambda: "0" * 16)
    entry, why, h = bench.ncu_constants("C2")
    assert entry is None and "stale" in why and h == "0" * 16
    # with a SASS table the digest of THIS workload's kernels decides, whatever the sources say:
    # an edit that leaves the workload's SASS alone keeps its capture, any other digest loses it
    monkeypatch.setattr(srchash, "kernel_sass_hashes", lambda lib_path=None: {"C2": table["C2"]["kernel_sass"]})
    entry, why, _ = bench.ncu_constants("C2")
    assert entry is not None
    monkeypatch.setattr(srchash, "kernel_sass_hashes", lambda lib_path=None: {"C2": "f" * 16})
    entry, why, _ = bench.ncu_constants("C2")
    assert entry is None and "SASS" in why
    entry, why, _ = bench.ncu_constants("C9")
    assert entry is None and "no capture" in why


def test_sass_table_is_bound_to_the_library_file(tmp_path, monkeypatch):
    import srchash

    side = tmp_path / "sass_hashes.json"
    lib = tmp_path / "lib.so"
    lib.write_bytes(b"one build")
    monkeypatch.setattr(srchash, "SIDECAR", str(side))
    side.write_text(json.dumps({"library_sha256": srchash._file_digest(str(lib)), "sass": {"C2": "abc"}}))
    assert srchash.kernel_sass_hashes(str(lib)) == {"C2": "abc"}
    lib.write_bytes(b"another build")
    assert srchash.kernel_sass_hashes(str(lib)) == {}


def test_registration_tool_sums_the_launches_of_a_step(tmp_path):
    raw = tmp_path / "raw.csv"
    hdr = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread"]
    units = ["", "", "ms", "Gbyte", "Mbyte", "inst", "%", "%", "%", "register/thread"]
    rows = [["0", "void azp::row_kernel<A, 0, 1>(...)", "1.5", "2.0", "100", "1000", "80", "40", "60", "100"],
            ["1", "void azp::row_kernel<A, 1, 0>(...)", "0.5", "0.5", "50", "500", "30", "20", "50", "110"],
            ["2", "void other_kernel(...)", "9", "9", "9", "9", "9", "9", "9", "9"]]
    with open(raw, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(hdr)
        w.writerow(units)
        w.writerows(rows)
    # run the tool on a copy of the repo layout so the committed table is not touched
    tools = tmp_path / "tools"
    tools.mkdir()
    (tmp_path / "profiles").mkdir()
    src = open(os.path.join(ROOT, "tools", "ncu_summary.py")).read()
    (tools / "ncu_summary.py").write_text(src)
    out = tmp_path / "profiles" / "s.csv"
    subprocess.run([sys.executable, str(tools / "ncu_summary.py"), str(raw), "--register", "CX", "--n", "7",
                    "--summary", str(out), "--hash", "abc", "--sass-hash", "def"], check=True, capture_output=True)
    entry = json.load(open(tmp_path / "profiles" / "ncu_constants.json"))["CX"]
    assert entry["kernel_sources"] == "abc" and entry["kernel_sass"] == "def"
    assert entry["N"] == 7 and len(entry["kernels"]) == 2
    assert entry["dram_bytes"] == 2.5e9 + 150e6
    assert entry["warp_instructions"] == 1500
    assert abs(entry["gpu_time_us_under_ncu"] - 2000.0) < 1e-6  # ms -> us
    assert entry["registers"] == 100  # of the longest launch


def test_sass_digest_does_not_depend_on_where_the_tree_is_checked_out(tmp_path, monkeypatch):
    """cuobjdump names the source file of an object by its absolute path ("identifier = ...");
    the digest ignores that line, so the library built from another checkout (or rebuilt on the
    GPU box) still matches the registered captures."""
    import shutil

    import srchash

    if not (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump")):
        import pytest
        pytest.skip("no cuobjdump")
    real_run = subprocess.run
    seen = []

    def fake_run(cmd, **kw):
        out = real_run(cmd, **kw)
        if "-sass" in cmd:
            seen.append(cmd[-1])
            assert b"identifier = " in out.stdout
            out.stdout = out.stdout.replace(b"identifier = /", b"identifier = /some/other/checkout/")
        return out

    side = tmp_path / "sass_hashes.json"
    monkeypatch.setattr(srchash, "SIDECAR", str(side))
    monkeypatch.setattr(srchash, "OBJECTS", {"C1": ["inst_plj.o"]})  # one object: seconds
    a = srchash.write_sass_hashes()
    if a is None:
        import pytest
        pytest.skip("library objects not built")
    monkeypatch.setattr(srchash.subprocess, "run", fake_run)
    b = srchash.write_sass_hashes()
    assert seen and a["sass"] == b["sass"]


def test_registered_captures_belong_to_the_kernels_of_this_build():
    """The committed ncu captures of C2-C5 were taken with exactly the SASS this tree builds
    (build() has run before the tests): bench.py will report their DRAM traffic."""
    import srchash

    here = srchash.kernel_sass_hashes()
    if not here:
        import pytest
        pytest.skip("no SASS table next to the library (build() not run)")
    table = json.load(open(os.path.join(ROOT, "profiles", "ncu_constants.json")))
    for wl in ("C2", "C3", "C4", "C5"):
        assert table[wl]["kernel_sass"] == here[wl], wl
