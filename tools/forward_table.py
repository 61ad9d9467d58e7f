#!/usr/bin/env python
"""ncu launch list of tools/one_forward.py (metrics gpu__time_duration.sum, sm__cycles_active.sum,
smsp__inst_executed.sum) -> per-kernel table of the LAST forward: time, SM-ms (sum over SMs of active cycles / clock)
and issue-ms (warp instructions / (4 per cycle and SM x 148 SMs) / clock: the share of the whole GPU's issue capacity).
    python tools/forward_table.py gpurun_out/forward.csv <launches per forward> > profiles/r2_forward_table.csv"""
import collections
import csv
import sys

CLK = 1.965e9


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    per_fwd = int(sys.argv[2]) if len(sys.argv) > 2 else None
    h = rows[0]
    ii, ik, im, iv, ig, ib = (h.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Grid Size", "Block Size"))
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(int(r[ii]), {"name": r[ik], "grid": r[ig], "block": r[ib]})
        d[r[im]] = float(r[iv].replace(",", ""))
    ids = sorted(launches)
    if per_fwd:
        ids = ids[-per_fwd:]
    per = collections.OrderedDict()
    for i in ids:
        d = launches[i]
        name = d["name"].split("(")[0].replace("void ", "")[:70]
        a = per.setdefault(name, [0, 0.0, 0.0, 0.0, d["grid"], d["block"]])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("sm__cycles_active.sum", 0.0)
        a[3] += d.get("smsp__inst_executed.sum", 0.0)
    tt = sum(v[1] for v in per.values())
    ts = sum(v[2] for v in per.values())
    ti = sum(v[3] for v in per.values())
    print("# last forward: %d launches, %.3f ms serialised, %.2f SM-ms active, %.3g warp instructions (%.3f ms of the whole "
          "GPU's issue capacity)" % (len(ids), tt / 1e6, ts / CLK * 1e3, ti, ti / (4 * 148) / CLK * 1e3))
    print("kernel,launches,first_grid,first_block,time_us,sm_ms_active,warp_inst,issue_ms_whole_gpu")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][2]):
        print("%s,%d,%s,%s,%.1f,%.3f,%.4g,%.4f" % (k.replace(",", ";"), v[0], v[4].replace(",", " "), v[5].replace(",", " "),
                                                  v[1] / 1e3, v[2] / CLK * 1e3, v[3], v[3] / (4 * 148) / CLK * 1e3))


if __name__ == "__main__":
    main()
