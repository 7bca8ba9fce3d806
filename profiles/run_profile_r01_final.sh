#!/bin/bash
# Round-1 final profiling recipe (run under gpurun, one GPU).  Each ncu pass follows a plain run of the SAME
# command that exited 0 (B200_PROFILING.md).  Raw outputs land in gpurun_out/; summaries are in profiles/
# (python profiles/summarize_ncu.py gpurun_out/r01_final_*.ncu-rep > profiles/r01_final_ncu.md).
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
timeout 300 $CMD > gpurun_out/final_plain_a.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01_final_launches.csv $CMD > gpurun_out/final_ncu_a.log 2>&1
echo "launch-list rc=$?"
timeout 120 python profiles/ncu_assign.py > gpurun_out/final_plain_b.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_select -s 2 -c 1 -f -o gpurun_out/r01_final_assign python profiles/ncu_assign.py > gpurun_out/final_ncu_b.log 2>&1
echo "assign capture rc=$?"
timeout 120 python profiles/ncu_knn.py > gpurun_out/final_plain_c.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_select -s 3 -c 1 -f -o gpurun_out/r01_final_knn python profiles/ncu_knn.py > gpurun_out/final_ncu_c.log 2>&1
echo "knn capture rc=$?"
timeout 120 python profiles/microbench_membound.py > gpurun_out/final_plain_d.log 2>&1 &&
timeout 400 ncu --set full --clock-control none -k regex:"histogram_warp|histogram_kernel|accumulate_priv|prepare_planes|absmax|normalize_l2|rescore_top1|okapi_dense|row_sum" -c 60 -f -o gpurun_out/r01_final_membound python profiles/microbench_membound.py > gpurun_out/final_ncu_d.log 2>&1
echo "membound capture rc=$?"
timeout 120 python profiles/ncu_hist.py > gpurun_out/final_plain_e.log 2>&1 &&
timeout 300 ncu --set full --clock-control none -k regex:"histogram_csr|exclusive_scan" -s 6 -c 3 -f -o gpurun_out/r01_final_csr python profiles/ncu_hist.py > gpurun_out/final_ncu_e.log 2>&1
echo "csr capture rc=$?"
ls -la gpurun_out | tail -8
