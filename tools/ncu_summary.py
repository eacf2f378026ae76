"""Text summary of an .ncu-rep (the files under profiles/ come from this script).
Usage: python tools/ncu_summary.py report.ncu-rep [kernel-substring ...] > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]


def main():
    rep, subs = sys.argv[1], sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    print("# %s" % rep.split("/")[-1])
    for row in rows[2:]:
        name = row[ki]
        if subs and not any(s in name for s in subs):
            continue
        print("--- %s" % name[:100])
        for k in KEYS:
            if k in h:
                print("   %-66s %s %s" % (k, row[h.index(k)], units[h.index(k)]))
        st = []
        for i, n in enumerate(h):
            if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued"):
                try:
                    st.append((float(row[i]), n[len("smsp__pcsamp_warps_issue_stalled_"):]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        print("   stall samples: " + ", ".join("%s %.1f%%" % (n, 100 * v / tot) for v, n in sorted(st, reverse=True)[:6]))


if __name__ == "__main__":
    main()
