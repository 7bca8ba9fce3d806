#!/bin/bash
# Round-1 profiling recipe (run under gpurun, one GPU).  Each ncu pass follows a plain run of the
# SAME command that exited 0 (B200_PROFILING.md).  Outputs land in gpurun_out/ and the summaries are
# copied into profiles/ by hand.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu_a.log 2>&1
echo "launch-list rc=$?"
# 15 assign launches precede the kNN ones (fit 2 + warm 3 + timed 2 + e2e 1+2 + fit 5): capture the last
# assign kernel and the first kNN kernel with the full section set
$CMD > gpurun_out/plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_select -s 14 -c 2 -f -o gpurun_out/r01_gemm_select $CMD > gpurun_out/ncu_b.log 2>&1
echo "full-capture rc=$?"
ls -la gpurun_out
