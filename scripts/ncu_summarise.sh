#!/bin/bash
# .ncu-rep -> the text summary committed under profiles/: every raw metric of the one captured launch, as CSV.
# usage: scripts/ncu_summarise.sh gpurun_out/r02_attention2_final.ncu-rep profiles/r02_attention2_final.metrics.csv
ncu -i "$1" --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ('Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration', 'sm__pipe_tensor', 'sm__inst_executed_pipe_tensor', 'dram__bytes', 'dram__throughput',
        'l1tex__data_pipe_lsu_wavefronts', 'lts__t_bytes', 'lts__t_sector_hit_rate', 'sm__throughput', 'gpu__compute_memory_throughput', 'smsp__cycles_active',
        'sm__warps_active', 'launch__registers_per_thread', 'launch__shared_mem', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed',
        'sm__cycles_elapsed', 'sm__cycles_active', 'smsp__pipe_xu', 'sm__inst_executed_pipe_xu', 'smsp__issue_active', 'l1tex__data_bank_conflicts',
        'smsp__average_warp', 'smsp__warp_issue_stalled', 'sm__pipe_fma', 'sm__pipe_alu', 'sm__inst_executed_pipe_uniform', 'gpc__cycles_elapsed.max', 'smsp__inst_executed_pipe_tmem', 'sm__mem')
w = csv.writer(sys.stdout)
w.writerow(['metric', 'unit', 'value'])
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(k) for k in keep):
        w.writerow([h, u, v])
" > "$2"
wc -l "$2"
