"""Reduce an `ncu --page raw --csv` export to the columns the roofline discussion needs (one row per profiled launch)."""
import csv, sys
KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum', 'launch__shared_mem_per_block_dynamic']
rows = list(csv.reader(open(sys.argv[1], newline='')))
hdr = rows[0]
units = rows[1] if len(rows) > 1 and not rows[1][0].isdigit() else None
idx = [hdr.index(k) for k in KEEP if k in hdr]
w = csv.writer(sys.stdout)
w.writerow([hdr[i] for i in idx])
if units:
    w.writerow([units[i] for i in idx])
for r in rows[2 if units else 1:]:
    if len(r) >= len(hdr):
        w.writerow([r[i][:90] for i in idx])
