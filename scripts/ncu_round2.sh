#!/bin/bash
# Round-2 ncu evidence for the fill kernel (run under gpurun, one GPU):
#   1. steady-state DRAM traffic: application replay, no cache control, launches 240.. of a
#      rotating chain -> dram__bytes_write.sum per launch ~ the 38.5 MB the launch writes
#   2. --set full of the same steady-state launches (pipe utilisation, occupancy, stalls, source)
#   3. the launch list of a short bench run (kernel shares of a step)
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
M=dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_op_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size
for DT in FP32 FP16; do
  timeout 600 ncu --replay-mode application --cache-control none --clock-control none -k regex:fill_uniform_kernel \
      --launch-skip 240 --launch-count 8 --metrics $M --csv --log-file gpurun_out/r02_fill_steady_dram_$DT.csv \
      python scripts/fill_steady.py 256 $DT > gpurun_out/ncu_steady_$DT.log 2>&1
done
timeout 900 ncu --set full --replay-mode application --cache-control none --clock-control none --import-source on \
    -k regex:fill_uniform_kernel --launch-skip 240 --launch-count 2 -o gpurun_out/r02_prof_fill_uniform \
    python scripts/fill_steady.py 256 FP32 > gpurun_out/ncu_fill_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 20 --warmup 5 --no-loopback --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_*.csv
