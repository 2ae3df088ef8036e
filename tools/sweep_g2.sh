for o in "L=24" "L=32" "L=48" "L=64"; do echo "#### OPTS=$o"; OPTS=$o GROUP=2 python tools/shard_perf.py 18 1,8 2>&1 | grep -E "^==|accumulate|fixup_d|bucket_red"; done
