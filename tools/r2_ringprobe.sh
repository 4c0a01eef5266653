#!/bin/bash
# Counter-level comparison of the flat copy and the TMA-ring copy skeleton: per-launch DRAM bytes, L2 sector
# traffic and hit rate, DRAM utilisation.  Outputs under gpurun_out/.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_srcunit_tex.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sectors_srcnode_gpc.sum
tools/ubench/copy2d > gpurun_out/copy2d_out.txt 2>&1
tools/ubench/tma_ring > gpurun_out/tma_ring_out.txt 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/copy2d_ncu.csv tools/ubench/copy2d > /dev/null 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/tma_ring_ncu.csv tools/ubench/tma_ring > /dev/null 2>&1
wc -l gpurun_out/copy2d_ncu.csv gpurun_out/tma_ring_ncu.csv; cat gpurun_out/copy2d_out.txt; cat gpurun_out/tma_ring_out.txt
