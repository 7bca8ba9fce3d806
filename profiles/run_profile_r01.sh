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
# full section set for the two hot shapes, from minimal drivers (one plain run each, then ncu):
#   profiles/ncu_assign.py : split top-1, 1M x 4096 x 128 (C2)
#   profiles/ncu_knn.py    : coarse seeded top-32, 10k x 1M x 2048 (C3)
python profiles/ncu_assign.py > gpurun_out/plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_select -s 2 -c 1 -f -o gpurun_out/r01_assign_final python profiles/ncu_assign.py > gpurun_out/ncu_b.log 2>&1
echo "assign capture rc=$?"
python profiles/ncu_knn.py > gpurun_out/plain_c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_select -s 3 -c 1 -f -o gpurun_out/r01_knn_final python profiles/ncu_knn.py > gpurun_out/ncu_c.log 2>&1
echo "knn capture rc=$?"
ls -la gpurun_out
