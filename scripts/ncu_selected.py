"""ncu report -> the small per-launch CSV kept under profiles/ (metric,unit,launch0,launch1,...).
usage: python scripts/ncu_selected.py report.ncu-rep kernel-regex out.csv"""
import csv, subprocess, sys
rep, kre, out = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kre],
                                      capture_output=True, text=True).stdout.splitlines()))
h, u, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
want += [n for n in h if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("_per_issue_active.ratio") and "not_issued" not in n]
with open(out, "w") as f:
    f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(data))) + "\n")
    for n in want:
        if n in h:
            i = h.index(n)
            f.write(",".join([n, u[i]] + ['"%s"' % r[i] if "," in r[i] else r[i] for r in data]) + "\n")
print("wrote", out, len(data), "launches")
