#!/bin/bash
# round 2 (third session), call F: position check of the A/B/A runs of calls D and E -- the same two libraries in the order
# old, new, old (a middle run that is slower whatever it runs would show up here)
L=tml_image_editing_defense_b200/csrc/libtml_b200.so
B="python bench.py --mode diffusion --batch 8 --unet native --steps 4 --warmup 2 --no_cpu_baseline"
cp $L /tmp/new.so
cp gpurun_ab_old.so $L; timeout 300 $B > gpurun_out/r3f_old1.json 2>/dev/null
cp /tmp/new.so $L; timeout 300 $B > gpurun_out/r3f_new.json 2>/dev/null
cp gpurun_ab_old.so $L; timeout 300 $B > gpurun_out/r3f_old2.json 2>/dev/null
cp /tmp/new.so $L
for f in gpurun_out/r3f_*.json; do echo $f; cut -c1-200 $f; done
