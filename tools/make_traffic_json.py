#!/usr/bin/env python
"""ncu launch list (--metrics dram__bytes_read.sum,dram__bytes_write.sum --csv) of `tools/one_forward.py
--default-options` -> profiles/r2_<tag>_traffic.json: DRAM read+write bytes per launch of the kernels matching a
regex, averaged over the LAST forward, together with the sha256 of the kernel source (bench.py reports the figure only
while that hash still matches).
    python tools/make_traffic_json.py gpurun_out/sa_traffic.csv "sa_fused_pipe_kernel|sa_inline_kernel" sa_fused \
        sa_common.cuh,sa_fused.cu,sa_inline.cu 5"""
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, regex, tag, source, per_forward = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ii, ik, im, iv, iu = (h.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r in rows[1:]:
        if re.search(regex, r[ik]) and r[im].startswith("dram__bytes_"):
            d = per.setdefault(int(r[ii]), {"name": r[ik]})
            d[r[im]] = float(r[iv].replace(",", "")) * scale[r[iu]]
    ids = sorted(per)[-per_forward:]
    tot = [per[i].get("dram__bytes_read.sum", 0) + per[i].get("dram__bytes_write.sum", 0) for i in ids]
    h = hashlib.sha256()
    for src in source.split(","):                          # comma-separated: the kernels of one op may live in several files
        h.update(open(os.path.join(ROOT, "spacap3d_b200", "csrc", src), "rb").read())
    sha = h.hexdigest()
    out = {"kernel": regex, "launches": len(ids), "dram_bytes_per_launch": [int(t) for t in tot],
           "dram_bytes_per_launch_avg": int(sum(tot) / max(len(tot), 1)), "source": "csrc/" + source,
           "source_sha256": sha,
           "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python "
                  "tools/one_forward.py --default-options (8 scenes x 40k points, config 2), last forward"}
    dst = os.path.join(ROOT, "profiles", "r2_%s_traffic.json" % tag)
    json.dump(out, open(dst, "w"), indent=1)
    print(dst, out["dram_bytes_per_launch_avg"], out["dram_bytes_per_launch"])


if __name__ == "__main__":
    main()
