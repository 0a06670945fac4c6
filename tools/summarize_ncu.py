#!/usr/bin/env python
"""ncu report (.ncu-rep) -> compact markdown table of the metrics the roofline needs.

    python tools/summarize_ncu.py gpurun_out/prof_x.ncu-rep [out.md]
"""
import csv
import io
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instr"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("smsp__inst_executed_pipe_fma.sum", "FMA instr"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
]


def main(path, out=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print("no data in", path)
        return
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        name = re.sub(r"\(.*$", "", rec.get("Kernel Name", "?"))
        lines.append(f"### `{name}`  (launch id {rec.get('ID', '?')})")
        lines.append("| metric | value | unit |")
        lines.append("|---|---:|---|")
        u = dict(zip(hdr, units))
        for key, label in WANT:
            if key in rec and rec[key] != "":
                lines.append(f"| {label} (`{key}`) | {rec[key]} | {u.get(key, '')} |")
        lines.append("")
    txt = "\n".join(lines)
    print(txt)
    if out:
        open(out, "w").write(txt + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
