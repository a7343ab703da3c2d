"""Identity of the kernels a measurement belongs to.

kernel_source_hash(): sha256 over the kernel sources that determine the pair-force SASS
(csrc/*.cuh, evaluators, inst_*.cu, capi.cu, Makefile) -- conservative: any edit changes it.

kernel_sass_hashes(): per workload, sha256 over the SASS (`cuobjdump -sass`) of the object files
that hold the kernels the workload launches (C2: inst_yukawa.o, C3: inst_colloid.o +
inst_hertz.o, ...), so a change that only touches one evaluator's kernels leaves the other
workloads' captures valid. Computed at build time (__graft_entry__.build) and stored next to
the library in azplugins_b200/sass_hashes.json together with the sha256 of libazp_b200.so; a
reader uses the table only if the library it would load still has that digest.

bench.py reports ncu-derived constants (DRAM traffic, warp instructions) only when the capture
they come from was taken with the same kernels: same SASS hash of the workload when both sides
have one, else the same source hash (profiles/ncu_constants.json records both at capture time).

    python tools/srchash.py            -> source hash
    python tools/srchash.py --sass C3  -> SASS hash of the workload's kernels ("" if unknown)
    python tools/srchash.py --write    -> (re)compute azplugins_b200/sass_hashes.json
"""
import glob
import hashlib
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJECTS = {"C1": ["inst_plj.o"], "C2": ["inst_yukawa.o"], "C3": ["inst_colloid.o", "inst_hertz.o"],
           "C4": ["inst_dpd.o"], "C5": ["inst_morse.o"]}
LIB = os.path.join(ROOT, "azplugins_b200", "libazp_b200.so")
SIDECAR = os.path.join(ROOT, "azplugins_b200", "sass_hashes.json")


def kernel_source_hash():
    base = os.path.join(ROOT, "azplugins_b200", "csrc")
    files = sorted(glob.glob(os.path.join(base, "*.cuh")) + glob.glob(os.path.join(base, "evaluators", "*.cuh"))
                   + glob.glob(os.path.join(base, "inst_*.cu")) + [os.path.join(base, "capi.cu"),
                                                                   os.path.join(base, "Makefile")])
    h = hashlib.sha256()
    for f in files:
        h.update(os.path.relpath(f, base).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def _file_digest(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 22), b""):
            h.update(chunk)
    return h.hexdigest()[:16]


def write_sass_hashes():
    """cuobjdump -sass over the workload objects of the current build -> sidecar. Returns the
    table, or None when cuobjdump or the objects are not there."""
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    build = os.path.join(ROOT, "azplugins_b200", "csrc", "build")
    if not os.path.exists(tool) or not os.path.exists(LIB):
        return None
    table = {}
    for wl, objs in OBJECTS.items():
        h = hashlib.sha256()
        for o in objs:
            path = os.path.join(build, o)
            if not os.path.exists(path):
                return None
            out = subprocess.run([tool, "-sass", path], capture_output=True, check=True).stdout
            # the dump names the source file of the object by its absolute path ("identifier =
            # /root/repo/..."): dropped, so the same SASS built in another checkout hashes the same
            out = b"\n".join(l for l in out.split(b"\n") if not l.startswith(b"identifier = "))
            h.update(o.encode())
            h.update(out)
        table[wl] = h.hexdigest()[:16]
    doc = {"library_sha256": _file_digest(LIB), "kernel_sources": kernel_source_hash(), "sass": table}
    json.dump(doc, open(SIDECAR, "w"), indent=1, sort_keys=True)
    return doc


def kernel_sass_hashes(lib_path=None):
    """{workload: hash} of the build the library at `lib_path` (default: the in-tree one) comes
    from, or {} when there is no sidecar for exactly that library file."""
    lib_path = lib_path or os.environ.get("AZP_B200_LIB") or LIB
    if not os.path.exists(SIDECAR) or not os.path.exists(lib_path):
        return {}
    doc = json.load(open(SIDECAR))
    if doc.get("library_sha256") != _file_digest(lib_path):
        return {}
    return doc.get("sass", {})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--write":
        print(json.dumps(write_sass_hashes()))
    elif len(sys.argv) > 2 and sys.argv[1] == "--sass":
        print(kernel_sass_hashes().get(sys.argv[2], ""))
    else:
        print(kernel_source_hash())
