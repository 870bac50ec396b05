# Round-2 call O: micro-batch 64 and two streams vs the default (same box)
B="python bench.py --quick --steps 5 --warmup 3"
timeout 300 $B 2>/dev/null | cut -c1-160
timeout 300 $B --micro_batch 64 2>/dev/null | cut -c1-160
timeout 300 $B --streams 2 2>/dev/null | cut -c1-160
timeout 300 $B --micro_batch 16 --streams 2 2>/dev/null | cut -c1-160
timeout 300 $B 2>/dev/null | cut -c1-160
