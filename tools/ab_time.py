#!/usr/bin/env python
"""A/B timing of library builds on the GPU box: for every library file given (variants built with
different kernel sources / macros, loaded through AZP_B200_LIB) and every workload, run the
bench's device-resident timing (autotuned launch shape, 20 steps) in a fresh process and print
ms/step.   tools/ab_time.py --libs a.so b.so --workloads C2 C4:2000000 C5:4000000"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", nargs="+", required=True)
    ap.add_argument("--workloads", nargs="+", default=["C2"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rows = []
    for wl in a.workloads:
        name, _, n = wl.partition(":")
        for lib in a.libs:
            env = dict(os.environ, AZP_B200_LIB=os.path.abspath(lib))
            cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", name, "--steps",
                   str(a.steps), "--warmup", "5", "--no-cpu-baseline", "--no-strong", "--no-e2e"]
            if n:
                cmd += ["--n-per-gpu", n]
            r = subprocess.run(cmd, env=env, capture_output=True, text=True)
            try:
                line = json.loads(r.stdout.strip().splitlines()[-1])
                row = dict(workload=wl, lib=os.path.basename(lib), ms=line["ms_per_step"],
                           shape=line["config"]["launch_shape_block_tpp_ms"],
                           frac=line["roofline"]["frac"], fusion=line["config"].get("fusion"))
            except Exception as exc:
                row = dict(workload=wl, lib=os.path.basename(lib), error=str(exc), stderr=r.stderr[-400:])
            rows.append(row)
            print(json.dumps(row), flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
