#!/bin/bash
# round 2, call W: correctness + launch durations of the fused attention kernels (A/B of kernel variants)
timeout 300 python tests/gpu_check_unet.py --which tiny > gpurun_out/w_tiny.log 2>&1; tail -2 gpurun_out/w_tiny.log
B="python bench.py --mode diffusion --unet native --batch 8 --steps 1 --warmup 1"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mh_attn_fwd|attn_bwd" --csv --log-file gpurun_out/w_attn_launches.csv $B > gpurun_out/w_a.log 2>&1
tail -1 gpurun_out/w_a.log | cut -c1-200
