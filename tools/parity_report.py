#!/usr/bin/env python
"""Print / save the model-level parity report of tests/refparity.py (stack A = unmodified reference on its own
CUDA extension, B = reference models on this package, C = spacap3d_b200.detector) at BASELINE size.

    python tools/parity_report.py [--dims 1 7 132] [--batch 8] [--points 40000] [--out gpurun_out/parity.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs="+", default=[1, 7, 132])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--points", type=int, default=40000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity.json"))
    args = ap.parse_args()
    import refparity
    reports = []
    for c in args.dims:
        r = refparity.collect(c, args.batch, args.points)
        reports.append(r)
        print("== C = %d (%s), %d x %d points" % (c, r["checkpoint"], args.batch, args.points))
        for tag in ("B_exact_vs_A", "B_fast_vs_A", "C_fast_vs_A"):
            print("  --", tag)
            for k, v in r[tag].items():
                print("     %-58s %s" % (k, ("%.3e" % v) if isinstance(v, float) else v))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(reports, f, indent=1)


if __name__ == "__main__":
    main()
