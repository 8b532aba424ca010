#!/bin/bash
# ncu --set full capture of a command, exported as a compact CSV of the metrics DESIGN.md / bench.py cite; the .ncu-rep is deleted
# (gpurun_out/ is capped at 64 MiB).  usage: ncu_export.sh <out.csv> <kernel regex> <count> <command...>
out=$1; re=$2; cnt=$3; shift 3
rep=/tmp/ncu_$$.ncu-rep
ncu --set full --clock-control none -k regex:$re -c $cnt -f -o ${rep%.ncu-rep} "$@" > /dev/null 2>&1
ncu -i $rep --page raw --csv 2>/dev/null | python3 -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
h = rows[0]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_uniform.sum', 'smsp__cycles_active.avg']
idx = [h.index(w) for w in want if w in h]
w = csv.writer(sys.stdout)
w.writerow([h[i] for i in idx])
w.writerow([rows[1][i] for i in idx])
for r in rows[2:]:
    w.writerow([r[i] for i in idx])
" > $out
rm -f $rep
