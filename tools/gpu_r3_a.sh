#!/bin/bash
# round 2 (third session), call A: stream / micro-batch sweep of the headline workload on one box (A/B inside one call)
Q="python bench.py --quick --steps 10 --warmup 3"
for cfg in "32 1" "32 2" "16 2" "16 4" "32 1" "32 2" "8 2"; do
  set -- $cfg
  echo "mb=$1 streams=$2: $(timeout 200 $Q --micro_batch $1 --streams $2 2>/dev/null | cut -c1-160)"
done | tee gpurun_out/r3a_streams.txt
