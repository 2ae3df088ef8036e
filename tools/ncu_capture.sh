#!/bin/bash
# ncu evidence for profiles/ (B200_PROFILING.md recipe): the launch list of one short bench run and one --set full
# capture of the hot kernels of a warm step.  Run under gpurun; the plain run comes first and must exit 0.
set -e
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --g2-logn 0 --groth16-logn 0"
TAG=${1:-r2}
$CMD > gpurun_out/${TAG}_ncu_plain.json 2> gpurun_out/${TAG}_ncu_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:BatchedAddRound|bucket_acc_kernel|bucket_reduce_kernel|row_sum_kernel' -s 24 -c 8 \
    -f -o gpurun_out/${TAG}_prof_hot $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
