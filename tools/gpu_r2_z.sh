#!/bin/bash
# round 2, call Z: final records of the native UNet path (bench lines, launch list, ncu --set full of the fused attention kernels)
B="python bench.py --mode diffusion --batch 8"
timeout 400 $B --unet native --steps 5 --warmup 2 --gemm_table > gpurun_out/z_native_b8.json 2> gpurun_out/z_native_b8_table.txt
timeout 400 $B --unet torch --steps 5 --warmup 2 > gpurun_out/z_torch_b8.json 2> /dev/null
timeout 400 $B --unet native --unet_recompute --steps 3 --warmup 1 > gpurun_out/z_native_b8_recompute.json 2> /dev/null
timeout 500 python bench.py --mode diffusion --batch 16 --unet native --steps 3 --warmup 1 > gpurun_out/z_native_b16.json 2> /dev/null
for k in mh_attn_fwd_online attn_bwd_dq attn_bwd_dkv; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/z_$k -f $B --unet native --steps 1 --warmup 0 > gpurun_out/z_$k.log 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/z_launches.csv $B --unet native --steps 1 --warmup 0 > gpurun_out/z_ncu.log 2>&1
for f in gpurun_out/z_*.json; do echo $f; cut -c1-230 $f; done
