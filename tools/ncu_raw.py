"""Key raw metrics of the first kernel in an ncu report: python tools/ncu_raw.py REPORT.ncu-rep [ROW]"""
import csv, subprocess, sys
rep = sys.argv[1]; row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); h = rows[0]; r = rows[2 + row]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__shared_mem_per_block_dynamic',
        'dram__bytes_read.sum', 'dram__bytes_write.sum']
keys += [k for k in h if 'issue_stalled' in k and k.endswith('per_issue_active.ratio')]
for k in keys:
    if k in h:
        v = r[h.index(k)]
        try:
            if 'stalled' in k and float(v) < 0.25: continue
        except ValueError: pass
        print(k.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', ''), v)
