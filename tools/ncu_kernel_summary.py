"""`ncu --set full` of ONE launch of a kernel, reduced to the metrics the design notes quote:

    python tools/ncu_kernel_summary.py <kernel-regex> <launch-skip> <out.txt> -- <command ...>

Profiling aid (GPU box)."""
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_red.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")


def main():
    regex, skip, out = sys.argv[1], sys.argv[2], sys.argv[3]
    cmd = sys.argv[sys.argv.index("--") + 1:]
    raw = out + ".csv"
    subprocess.check_call(["ncu", "--set", "full", "--clock-control", "none", "-k", "regex:" + regex, "--launch-skip", skip,
                           "--launch-count", "1", "--csv", "--page", "raw", "--log-file", raw] + cmd,
                          stdout=subprocess.DEVNULL)
    rows = list(csv.reader(ln for ln in open(raw) if not ln.startswith("==")))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(out, "w") as f:
        f.write("== %s (launch-skip %s): %s\n" % (regex, skip, " ".join(cmd)))
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP or h == "Kernel Name" or "pcsamp_warps_issue_stalled" in h:
                if "pcsamp" in h and (not v or float(v.replace(",", "") or 0) < 500):
                    continue
                f.write("  %-80s %s %s\n" % (h, v, u))


if __name__ == "__main__":
    main()
