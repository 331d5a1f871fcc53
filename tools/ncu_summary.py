#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep from `ncu --set full`) or a launch list csv into a markdown table under profiles/.

    python tools/ncu_summary.py full  gpurun_out/prof.ncu-rep  profiles/rNN_name.md  "title / command line"
    python tools/ncu_summary.py list  gpurun_out/launches.csv  profiles/rNN_name.md  "title / command line"
"""
import collections
import csv
import io
import re
import subprocess
import sys

FULL = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
SHORT = ["time us", "dram rd MB", "dram wr MB", "regs", "grid", "warps act %", "SM thr %", "L1 thr %", "DRAM thr %", "tensor %", "warp insts"]


def short_name(n):
    n = n.replace("void ", "").replace("cotb200::", "")
    n = re.sub(r"\(.*", "", n)
    return n[:90]


def full(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    ix = [hdr.index(m) if m in hdr else None for m in FULL]
    lines = ["# " + title, "", "| kernel | " + " | ".join(SHORT) + " |", "|---|" + "---|" * len(SHORT)]
    traffic = {}
    for r in rows[2:]:
        vals = []
        for i, m in zip(ix, FULL):
            if i is None or r[i] == "":
                vals.append("-")
                continue
            v = float(r[i].replace(",", ""))
            u = units[i]
            if m.startswith("dram__bytes"):
                scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                v *= scale
                vals.append("%.1f" % v)
            elif m == "gpu__time_duration.sum":
                scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                vals.append("%.1f" % (v * scale))
            elif m == "smsp__inst_executed.sum":
                vals.append("%.3g" % v)
            else:
                vals.append("%.0f" % v)
        lines.append("| `%s` | %s |" % (short_name(r[kn]), " | ".join(vals)))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def launch_list(path, out, title, top=40):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    mu = hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for r in rows[start:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        t = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1e-3)
        name = re.sub(r"<.*", "", re.sub(r"\(.*", "", r[kn]))[:80]
        agg[name][0] += 1
        agg[name][1] += t
        n += 1
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if "cotb200" in k)
    lines = ["# " + title, "", "%d launches, %.2f ms of kernel time (cold-cache, serialised by ncu: compare SHARES); "
             "libcotb200 kernels: %.2f ms = %.1f %%" % (n, tot / 1e3, ours / 1e3, 100 * ours / tot), "",
             "| ms | share % | launches | kernel |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        lines.append("| %.2f | %.1f | %d | `%s` |" % (v[1] / 1e3, 100 * v[1] / tot, v[0], k))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:30]))


if __name__ == "__main__":
    mode, src, dst, title = sys.argv[1:5]
    (full if mode == "full" else launch_list)(src, dst, title)
