#!/bin/bash
# ncu --set full capture of the G2 hot kernels of one warm MSM at 2^18 (profiles/): plain run first, must exit 0
set -e
export GROUP=2
python tools/shard_perf.py 18 1 > gpurun_out/r2_g2_ncu_plain.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:BatchedAddRound|Accumulate|FixupDirect|bucket_reduce_kernel|row_sum_kernel' -s 60 -c 6 \
    -f -o gpurun_out/r2_prof_g2 python tools/shard_perf.py 18 1 > gpurun_out/r2_g2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_g2_ncu_full.log
