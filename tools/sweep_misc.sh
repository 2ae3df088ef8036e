for o in "" "min_left=6" "min_left=3"; do echo "#### 2^20 OPTS=$o"; OPTS=$o python tools/shard_perf.py 20 1 2>&1 | grep -E "^==|batched|accumulate_b"; done
for o in "" "min_left=6" "min_left=3" "coop_max_chains=40000" "min_left=6,coop_max_chains=40000"; do echo "#### 2^24 OPTS=$o"; OPTS=$o python tools/shard_perf.py 24 1 2>&1 | grep -E "^==|batched|accumulate_b|bucket_red|pair_sum|row_sum"; done
for o in "" "min_left=6"; do echo "#### 2^22 OPTS=$o"; OPTS=$o python tools/shard_perf.py 22 1 2>&1 | grep -E "^==|batched|accumulate_b|bucket_red"; done
