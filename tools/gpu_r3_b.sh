#!/bin/bash
# round 2 (third session), call B: validation of HEAD -- GPU test suite, smoke(), default bench line
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3b_test.log 2>&1; tail -2 gpurun_out/r3b_test.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3b_smoke.log 2>&1; tail -2 gpurun_out/r3b_smoke.log
timeout 600 python bench.py > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err; cut -c1-400 gpurun_out/r3b_bench.json
