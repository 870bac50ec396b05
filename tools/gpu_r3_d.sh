#!/bin/bash
# round 2 (third session), calls D, E: UNet elementwise kernels (D: batched loads / finer chunks in the general GroupNorm, 32-bit
# index math; E: four loads per thread in head split / merge / column copies, two rows per warp in LayerNorm, two vector pairs
# per thread in the GEGLU forward): tests, then A/B/A of the diffusion bench against the previous library (gpurun_ab_old.so)
L=tml_image_editing_defense_b200/csrc/libtml_b200.so
timeout 900 python -m pytest tests/test_gpu_unet.py -x -q > gpurun_out/r3e_test.log 2>&1; tail -2 gpurun_out/r3e_test.log
B="python bench.py --mode diffusion --batch 8 --unet native --steps 4 --warmup 2 --no_cpu_baseline"
cp $L /tmp/new.so
timeout 300 $B > gpurun_out/r3e_new1.json 2>/dev/null
cp gpurun_ab_old.so $L; timeout 300 $B > gpurun_out/r3e_old.json 2>/dev/null
cp /tmp/new.so $L; timeout 300 $B > gpurun_out/r3e_new2.json 2>/dev/null
for f in gpurun_out/r3e_*.json; do echo $f; cut -c1-200 $f; done
