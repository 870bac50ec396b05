#!/bin/bash
# round 2 (third session), call J: last check of the committed tree (GPU suite + smoke)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3j_test.log 2>&1; tail -2 gpurun_out/r3j_test.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3j_smoke.log 2>&1; tail -1 gpurun_out/r3j_smoke.log
