#!/usr/bin/env python
"""Condense gpurun_out/parity.json (tools/parity_report.py) into the table committed under profiles/."""
import json
import sys

R = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity.json"))
for r in R:
    print("C = %d  (%s, %d x %d points)" % (r["feature_dim"], r["checkpoint"], r["batch"], r["n_points"]))
    for tag in ("B_exact_vs_A", "B_fast_vs_A", "C_fast_vs_A"):
        d = r[tag]
        eq = [k for k, v in d.items() if k.endswith("equal") and "bit_equal" not in k]
        bad = [k for k in eq if d[k] is False]
        print("  %-13s index/coordinate/ball-query/decode equalities: %d of %d hold%s" % (
            tag, len(eq) - len(bad), len(eq), ("  NOT: " + ", ".join(bad)) if bad else ""))
        ne = {k[:-5]: v for k, v in d.items() if k.endswith(":nerr")}
        if tag == "B_exact_vs_A":
            print("      float tensors bit-equal: %s; max nerr %.1e" % (
                all(v for k, v in d.items() if k.endswith("bit_equal")), max(ne.values())))
        else:
            print("      nerr = max|x-ref|/(|ref|+rms):  " + "  ".join("%s %.1e" % (k, v) for k, v in ne.items()))
            print("      other: " + "  ".join("%s %.4g" % (k, v) for k, v in d.items()
                                              if "fraction" in k or "agree" in k or "maxabs" in k))
