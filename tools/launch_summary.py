#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum,sm__cycles_active.sum --csv --log-file X.csv) -> per-kernel
summary csv: launches, total time, time share, share of SM-active cycles.
    python tools/launch_summary.py gpurun_out/launches.csv "command that was profiled" > profiles/rN_launches_bench_summary.csv"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    cmd = sys.argv[2] if len(sys.argv) > 2 else "?"
    h = rows[0]
    ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    per = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        if r[im] == "gpu__time_duration.sum":
            per[r[ik]][0] += 1
            per[r[ik]][1] += v
        elif r[im] == "sm__cycles_active.sum":
            per[r[ik]][2] += v
    tt, tc, n = sum(v[1] for v in per.values()), sum(v[2] for v in per.values()), sum(v[0] for v in per.values())
    print("# ncu launch list of `%s` (%d launches)" % (cmd, n))
    print("# per-launch times are cold-cache and serialised: only each kernel's SHARE is meaningful (B200, --clock-control none)")
    print("# total %.2f ms; sm__cycles_active total %.3g (a kernel alone on the GPU: its CTAs spread over more SMs than "
          "when batches overlap)" % (tt / 1e6, tc))
    print("kernel,launches,time_ms,time_share,sm_cycles_active_share")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
        name = k.replace(",", ";")
        if len(name) > 90:
            name = name[:90]
        print("%s,%d,%.3f,%.4f,%.4f" % (name, v[0], v[1] / 1e6, v[1] / tt, v[2] / tc if tc else 0))


if __name__ == "__main__":
    main()
