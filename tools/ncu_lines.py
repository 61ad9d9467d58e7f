#!/usr/bin/env python
"""Aggregate an ncu source page (ncu -i rep --page source --csv --print-source cuda,sass > src.csv) per CUDA source
line: samples, warp instructions, top stall reasons.   python tools/ncu_lines.py src.csv [kernel_substring] [top]"""
import collections
import csv
import sys


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "File Path":
            cur = {"file": r[1], "rows": []}
            secs.append(cur)
        elif r and r[0] == "Function Name":
            cur["fn"] = r[1]
        elif r and r[0] == "Line No":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    seen = set()
    for s in secs:
        if want not in s.get("fn", "") or len(s["rows"]) < 2000 or s["fn"] in seen:
            continue
        seen.add(s["fn"])
        h = s["hdr"]
        iS, iI = h.index("# Samples"), h.index("Instructions Executed")
        stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), ""])
        for r in s["rows"]:
            if r[0] == "":
                continue
            a = agg[r[0]]
            a[0] += I(r[iS]); a[1] += I(r[iI]); a[3] = r[1]
            for i in stall:
                a[2][h[i]] += I(r[i])
        print("==", s["fn"][:110], "samples", sum(a[0] for a in agg.values()), "inst", sum(a[1] for a in agg.values()))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            print("%5s %6d %9d  %-78s %s" % (k, a[0], a[1], a[3][:78].strip(),
                                             " ".join("%s=%d" % (n[6:], c) for n, c in a[2].most_common(3))))


if __name__ == "__main__":
    main()
