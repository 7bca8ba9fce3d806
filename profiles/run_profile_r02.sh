#!/bin/bash
# Round-2 profiling recipe (run under gpurun, one GPU).  Each ncu pass follows a plain run of the SAME command that
# exited 0 (B200_PROFILING.md).  Raw outputs land in gpurun_out/; summaries are written by
#   python profiles/summarize_ncu.py --traffic-json profiles/r02_ncu_traffic.json gpurun_out/r02_*.ncu-rep > profiles/r02_ncu.md
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-c4 --no-c5"
timeout 300 $CMD > gpurun_out/r02_plain_a.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_a.log 2>&1
echo "launch-list rc=$?"
# four gemm_select launches per verified fused assign (one-product pass, compact split re-run, two gated launches that
# return at once): skip two calls, capture the third call's pass and re-run
timeout 120 python profiles/ncu_assign.py > gpurun_out/r02_plain_b.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_select -s 8 -c 2 -f -o gpurun_out/r02_assign python profiles/ncu_assign.py > gpurun_out/r02_ncu_b.log 2>&1
echo "assign capture rc=$?"
timeout 120 python profiles/ncu_knn.py > gpurun_out/r02_plain_c.log 2>&1 &&
timeout 400 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section LaunchStats --section Occupancy --section WarpStateStats --section ComputeWorkloadAnalysis --clock-control none -k regex:gemm_select -s 3 -c 1 -f -o gpurun_out/r02_knn python profiles/ncu_knn.py > gpurun_out/r02_ncu_c.log 2>&1
echo "knn capture rc=$?"
timeout 120 python profiles/ncu_r02_membound.py > gpurun_out/r02_plain_d.log 2>&1 &&
timeout 400 ncu --set full --clock-control none -k regex:"prepare_rows|gather_reduce|scatter_kernel|count_kernel|scan_counts" -s 5 -c 6 -f -o gpurun_out/r02_membound python profiles/ncu_r02_membound.py > gpurun_out/r02_ncu_d.log 2>&1
echo "membound capture rc=$?"
timeout 120 python profiles/ncu_hist.py > gpurun_out/r02_plain_e.log 2>&1 &&
timeout 300 ncu --set full --clock-control none -k regex:"histogram" -s 2 -c 2 -f -o gpurun_out/r02_hist python profiles/ncu_hist.py > gpurun_out/r02_ncu_e.log 2>&1
echo "hist capture rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches.csv
