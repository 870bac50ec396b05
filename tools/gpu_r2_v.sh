#!/bin/bash
# round 2, call V: launch durations of the fused attention kernels + one --set full capture of each (64 x 64 latents)
B="python bench.py --mode diffusion --unet native --batch 8 --steps 1 --warmup 1"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mh_attn_fwd|attn_bwd" --csv --log-file gpurun_out/v_attn_launches.csv $B > gpurun_out/v_a.log 2>&1
for k in mh_attn_fwd attn_bwd_dq attn_bwd_dkv; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/v_$k -f $B > gpurun_out/v_$k.log 2>&1
done
ls -la gpurun_out/v_*
