"""Write a text summary (key metrics + stall reasons) of the first kernels of an .ncu-rep:  python scripts/ncu_summary.py rep out.txt 'header line' ..."""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
hdrs = sys.argv[3:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg.per_second',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
lines = ["# " + h for h in hdrs]
for vals in rows[2:]:
    lines.append("")
    for i, h in enumerate(hdr):
        try:
            big = float(vals[i].replace(",", "") or 0) > 0.05
        except ValueError:
            big = False
        if h in keep or ('issue_stalled' in h and 'per_issue_active' in h and big):
            lines.append("%-92s %-16s %s" % (h, units[i], vals[i]))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
