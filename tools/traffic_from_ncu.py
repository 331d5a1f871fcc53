#!/usr/bin/env python
"""ncu --csv metric log (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch) ->
profiles/traffic.json {prof name: mean DRAM bytes per launch} + a markdown table.

    python tools/traffic_from_ncu.py gpurun_out/traffic_bench.csv profiles/traffic.json profiles/r01_traffic.md"""
import collections
import csv
import json
import re
import sys

NAMES = {"bn_bwd_apply_kernel": "bn_bwd_apply", "bn_bwd_sums_kernel": "bn_bwd_sums", "bn_apply_kernel": "bn_apply",
         "col_stats_kernel": "col_stats"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
        "nsecond": 1e-3, "second": 1e6}


def main():
    src, dst_json, dst_md = sys.argv[1:4]
    lines = [l for l in open(src) if not l.startswith("==")]
    per = collections.defaultdict(lambda: collections.defaultdict(dict))
    for row in csv.DictReader(lines):
        m = re.search(r"(\w+_kernel)\b", row["Kernel Name"])
        if not m or m.group(1) not in NAMES:
            continue
        val = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
        per[NAMES[m.group(1)]][row["ID"]][row["Metric Name"]] = val
    out, md = {}, ["# r01 — DRAM traffic of the BatchNorm kernels over whole bench steps (ncu, per-launch means)", "",
                   "| kernel | launches | DRAM read+write MB / launch | us / launch (ncu, cold) |", "|---|---|---|---|"]
    for k, launches in sorted(per.items()):
        tot = [v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0) for v in launches.values()]
        us = [v.get("gpu__time_duration.sum", 0.0) for v in launches.values()]
        out[k] = sum(tot) / len(tot)
        md.append("| %s | %d | %.1f | %.1f |" % (k, len(tot), out[k] / 1e6, sum(us) / len(us)))
    json.dump(out, open(dst_json, "w"), indent=1)
    open(dst_md, "w").write("\n".join(md) + "\n")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
