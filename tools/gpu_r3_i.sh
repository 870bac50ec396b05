#!/bin/bash
# round 2 (third session), call I: UNet tests + smoke after the LayerNorm forward went back to one row per warp
timeout 600 python -m pytest tests/test_gpu_unet.py -x -q > gpurun_out/r3i_test.log 2>&1; tail -2 gpurun_out/r3i_test.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3i_smoke.log 2>&1; tail -1 gpurun_out/r3i_smoke.log
