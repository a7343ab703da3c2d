"""sha256 over the kernel sources that determine the pair-force SASS (csrc/*.cuh, evaluators,
inst_*.cu, capi.cu, Makefile). bench.py reports ncu-derived constants (DRAM traffic, warp
instructions) only when the capture they come from was taken with the same sources
(profiles/ncu_constants.json records the hash at capture time)."""
import glob
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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


if __name__ == "__main__":
    print(kernel_source_hash())
