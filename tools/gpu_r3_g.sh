#!/bin/bash
# round 2 (third session), call G: validation of the final tree -- GPU test suite, smoke(), default bench line, diffusion record
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3g_test.log 2>&1; tail -2 gpurun_out/r3g_test.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3g_smoke.log 2>&1; tail -1 gpurun_out/r3g_smoke.log
timeout 600 python bench.py > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err; cut -c1-200 gpurun_out/r3g_bench.json
timeout 400 python bench.py --mode diffusion --batch 8 --unet native --steps 5 --warmup 2 --gemm_table > gpurun_out/r3g_unet_b8.json 2> gpurun_out/r3g_unet_b8_table.txt; cut -c1-200 gpurun_out/r3g_unet_b8.json
