#!/usr/bin/env python
"""Per-launch detail of an `ncu --set full` report: duration, DRAM / L2 / L1 / SM throughput, occupancy, tensor pipe, issue
activity and the top warp-stall reasons -- the numbers the kernel notes in DESIGN.md and profiles/*.md quote.

    python tools/ncu_detail.py gpurun_out/targets.ncu-rep [profiles/rNN_name.md] [--filter regex]
"""
import csv
import io
import re
import subprocess
import sys

COLS = [
    ("time us", "gpu__time_duration.sum", "time"),
    ("dram rd MB", "dram__bytes_read.sum", "bytes"),
    ("dram wr MB", "dram__bytes_write.sum", "bytes"),
    ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "pct"),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "pct"),
    ("L1 %", "l1tex__throughput.avg.pct_of_peak_sustained_active", "pct"),
    ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "pct"),
    ("issue %", "sm__inst_issued.avg.pct_of_peak_sustained_active", "pct"),
    ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "pct"),
    ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active", "pct"),
    ("regs", "launch__registers_per_thread", "int"),
    ("grid", "launch__grid_size", "int"),
    ("block", "launch__block_size", "int"),
    ("smem KB", "launch__shared_mem_per_block_dynamic", "kb"),
]


def short_name(n):
    n = n.replace("void ", "").replace("cotb200::", "")
    n = re.sub(r"\(.*", "", n)
    return n[:60]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    flt = None
    if "--filter" in sys.argv:
        flt = re.compile(sys.argv[sys.argv.index("--filter") + 1])
        args = [a for a in args if a != sys.argv[sys.argv.index("--filter") + 1]]
    rep = args[0]
    out = args[1] if len(args) > 1 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    stall_ix = [(i, h) for i, h in enumerate(hdr) if re.match(r"smsp__average_warps?_issue_stalled_.*_per_issue_active", h)
                or re.match(r"smsp__average_warp_latency_issue_stalled_.*\.ratio", h)]
    lines = ["| # | kernel | " + " | ".join(c[0] for c in COLS) + " | top stalls (warps per issue) |", "|---|---|" + "---|" * (len(COLS) + 1)]
    for n, r in enumerate(rows[2:]):
        name = short_name(r[kn])
        if flt and not flt.search(r[kn]):
            continue
        vals = []
        for _, m, kind in COLS:
            if m not in hdr or r[hdr.index(m)] == "":
                vals.append("-")
                continue
            i = hdr.index(m)
            v = float(r[i].replace(",", ""))
            u = units[i]
            if kind == "bytes":
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                vals.append("%.1f" % v)
            elif kind == "time":
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(u, 1.0)
                vals.append("%.1f" % v)
            elif kind == "kb":
                v *= {"byte": 1e-3, "Kbyte": 1.0, "Mbyte": 1e3}.get(u, 1e-3)
                vals.append("%.0f" % v)
            elif kind == "int":
                vals.append("%d" % v)
            else:
                vals.append("%.0f" % v)
        st = []
        for i, h in stall_ix:
            try:
                st.append((float(r[i].replace(",", "")), re.sub(r"smsp__average_warps?_(latency_)?issue_stalled_|_per_issue_active.*|\.ratio", "", h)))
            except ValueError:
                pass
        st.sort(reverse=True)
        lines.append("| %d | `%s` | %s | %s |" % (n, name, " | ".join(vals), ", ".join("%s %.1f" % (b, a) for a, b in st[:4])))
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write("# ncu --set full detail of %s\n\n%s\n" % (rep, text))


if __name__ == "__main__":
    main()
