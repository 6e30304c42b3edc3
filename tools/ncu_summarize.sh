#!/bin/bash
# Text summary of one .ncu-rep for profiles/: key lines of the details page, DRAM bytes / pipe use from the raw page, stall
# reasons and hottest SASS instructions from the source page.    tools/ncu_summarize.sh X.ncu-rep > profiles/X.txt
R=$1
ncu -i $R --page details 2>/dev/null | grep -E "^  [a-z_].*\(|SM Frequency|Elapsed Cycles|Memory Throughput|DRAM Throughput|Duration|Executed Ipc Active|Issue Slots Busy|L2 Hit Rate|Eligible Warps|Warp Cycles Per Issued|Executed Instructions|Cluster Size|Grid Size|Registers Per Thread|Dynamic Shared Memory Per Block|Max Active Clusters|Achieved Active Warps|Achieved Occupancy|Theoretical Occupancy"
ncu -i $R --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr,units=rows[0],rows[1]
want=['dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    print('launch:', r[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '')
    for w in want:
        if w in hdr: print(f'  {w:90s} {units[hdr.index(w)]:>12s} {r[hdr.index(w)]:>16s}')
"
ncu -i $R --page source --csv 2>/dev/null > /tmp/ncu_src.csv && python3 $(dirname $0)/ncu_hotspots.py /tmp/ncu_src.csv 14
