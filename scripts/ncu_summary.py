"""Key metrics of one kernel from an ncu report. usage: python scripts/ncu_summary.py report.ncu-rep [kernel-regex]"""
import csv, subprocess, sys
rep = sys.argv[1]
cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"]
if len(sys.argv) > 2: cmd += ["--kernel-name", "regex:" + sys.argv[2]]
rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
h, u = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__average_warps_issue_stalled', 'sm__inst_executed_pipe', 'smsp__sass_average_branch_targets_threads_uniform.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__inst_executed_op_local', 'sm__ctas_launched']
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:100])
    for i, n in enumerate(h):
        if any(n.startswith(w) for w in want) and not n.endswith(("per_second", "peak_sustained_elapsed")) or n == 'dram__throughput.avg.pct_of_peak_sustained_elapsed':
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if v != 0: print(f"  {n:95s} {u[i]:>14s} {r[i]}")
