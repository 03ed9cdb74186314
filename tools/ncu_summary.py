#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into the text kept under profiles/."""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of {rep} (caches flushed between replays: DRAM bytes are cold-cache)\n")
        for d in data:
            f.write("\n== " + d[hdr.index("Kernel Name")] + "\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"  {k:72s} {d[i]} {units[i]}\n")
            st = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                    try: st.append((float(d[i].replace(",", "")), h))
                    except ValueError: pass
            st.sort(reverse=True)
            f.write("  top stalls (warps per issue-active cycle): " + ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for v, h in st[:6]) + "\n")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
