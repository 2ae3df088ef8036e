for o in "" "batch_blocks=2" "batch_blocks=1"; do echo "#### OPTS=$o"; OPTS=$o python tools/shard_perf.py 20 1,8 2>&1 | grep -E "^==|batched"; done
