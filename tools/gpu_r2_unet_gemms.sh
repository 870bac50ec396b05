#!/bin/bash
# round 2: the UNet's 3x3 convolution shapes in isolation (16 samples), and an ncu --set full capture of the 64 x 64 case
M="python tools/gemm_micro.py --bias --iters 20"
$M --B 16 --H 64 --W 64 --Cin 320 --N 320 --tag "64^2 320->320 (h66, BN=160)"
TML_H66_PAIR_MIN_BN=160 $M --B 16 --H 64 --W 64 --Cin 320 --N 320 --tag "64^2 320->320 (h66 pairs)"
$M --B 16 --H 64 --W 64 --Cin 640 --N 320 --tag "64^2 640->320"
$M --B 16 --H 32 --W 32 --Cin 640 --N 640 --tag "32^2 640->640"
$M --B 16 --H 32 --W 32 --Cin 1280 --N 640 --tag "32^2 1280->640"
$M --B 16 --H 16 --W 16 --Cin 1280 --N 1280 --tag "16^2 1280->1280"
TML_NO_BN_HEUR=1 $M --B 16 --H 16 --W 16 --Cin 1280 --N 1280 --tag "16^2 1280->1280 (BN=256 pairs)"
$M --B 16 --H 8 --W 8 --Cin 1280 --N 1280 --tag "8^2 1280->1280"
TML_NO_BN_HEUR=1 $M --B 16 --H 8 --W 8 --Cin 1280 --N 1280 --tag "8^2 1280->1280 (BN=256)"
$M --B 32 --H 8 --W 8 --Cin 1280 --N 1280 --tag "8^2 1280->1280, 32 samples"
timeout 300 ncu --set full --clock-control none -k regex:conv_gemm_tcgen05 -c 1 --launch-skip 3 -o gpurun_out/unet_conv64 -f $M --B 16 --H 64 --W 64 --Cin 320 --N 320 > gpurun_out/unet_conv64.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:conv_gemm_tcgen05 -c 1 --launch-skip 3 -o gpurun_out/unet_conv8 -f $M --B 16 --H 8 --W 8 --Cin 1280 --N 1280 > gpurun_out/unet_conv8.log 2>&1
