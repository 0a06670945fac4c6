#!/usr/bin/env python
"""ncu --set full reports -> per-kernel-class DRAM traffic table (profiles/r2_traffic.{json,md}) for bench.py's
`roofline.traffic`:  python tools/traffic_table.py gpurun_out/r2i_*.ncu-rep"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"dur_us": "gpu__time_duration.sum", "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum",
        "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l2_hit": "lts__t_sector_hit_rate.pct", "grid": "launch__grid_size"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "ns": 1e-3, "us": 1, "ms": 1e3}


def rows(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--kernel-name-base", "demangled"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    if len(rd) < 3:
        return
    hdr, units = rd[0], rd[1]
    for r in rd[2:]:
        rec, un = dict(zip(hdr, r)), dict(zip(hdr, units))
        name = re.sub(r"\(int\)", "", rec.get("Kernel Name", "?"))
        name = re.sub(r"\(mtts::.*$", "", name).replace("void mtts::", "").replace("mtts::", "")
        out = {"kernel": name}
        for k, m in KEYS.items():
            if m in rec and rec[m] != "":
                out[k] = float(rec[m].replace(",", "")) * UNIT.get(un.get(m, ""), 1)
        yield out


def main(paths):
    agg = collections.OrderedDict()
    for p in paths:
        for r in rows(p):
            a = agg.setdefault(r["kernel"], collections.defaultdict(list))
            for k, v in r.items():
                if k != "kernel":
                    a[k].append(v)
    table = []
    for name, a in agg.items():
        n = len(a["dur_us"])
        mean = lambda k: sum(a[k]) / len(a[k]) if a[k] else None
        table.append({"kernel": name, "launches_captured": n, "avg_duration_us": round(mean("dur_us"), 1),
                      "dram_bytes_per_launch": int(mean("dram_rd") + mean("dram_wr")), "dram_read": int(mean("dram_rd")),
                      "dram_write": int(mean("dram_wr")), "dram_pct_of_peak": round(mean("dram_pct"), 1),
                      "tensor_pipe_active_pct": round(mean("tensor_pct"), 1) if a["tensor_pct"] else None,
                      "l2_hit_pct": round(mean("l2_hit"), 1) if a["l2_hit"] else None})
    dom = next((t for t in table if t["kernel"].startswith("conv_tc_kernel<128, 128, 1")), table[0] if table else None)
    out = {"how": "ncu --set full --clock-control none (one replayed capture per launch; cold-ish caches), gpurun call I; averages over the "
                  "captured launches of each template instance at batch 64",
           "dominant_kernel": dom["kernel"] if dom else None,
           "dominant_kernel_dram_bytes_per_launch": dom["dram_bytes_per_launch"] if dom else None,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the CTA-pair dense-layer instance (PLM late steps, "
                   "M ~ 3.5-4k rows); per-class table in profiles/r2_traffic.md",
           "classes": table}
    with open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    lines = ["# Measured DRAM traffic per launch of the top kernel classes (ncu --set full, batch 64)", "",
             "| kernel | captured | avg us | DRAM read | DRAM write | DRAM % of peak | tensor pipe active % | L2 hit % |",
             "|---|---:|---:|---:|---:|---:|---:|---:|"]
    for t in table:
        lines.append(f"| `{t['kernel']}` | {t['launches_captured']} | {t['avg_duration_us']} | {t['dram_read'] / 1e6:.1f} MB | "
                     f"{t['dram_write'] / 1e6:.1f} MB | {t['dram_pct_of_peak']} | {t['tensor_pipe_active_pct']} | {t['l2_hit_pct']} |")
    with open(os.path.join(ROOT, "profiles", "r2_traffic.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1:])
