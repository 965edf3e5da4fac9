#!/usr/bin/env python
"""Summarise an ncu report (read on the GPU-less build box) into a small CSV/markdown for profiles/.
usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_xxx"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_ncu_peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["| kernel | " + " | ".join(n for _, n in METRICS) + " | top stalls (pc samples) |", "|" + "---|" * (len(METRICS) + 2)]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        vals = []
        for m, _ in METRICS:
            vals.append(f"{r[idx[m]]} {units[idx[m]]}".strip() if m in idx else "n/a")
        st = []
        for h, i in idx.items():
            if h.startswith(STALLS) and not h.endswith("_not_issued"):
                try:
                    st.append((float(r[i].replace(",", "")), h[len(STALLS):]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        top = ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:4])
        lines.append(f"| {name} | " + " | ".join(vals) + f" | {top} |")
    open(out + ".md", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
