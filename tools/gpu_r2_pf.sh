#!/bin/bash
# round 2: L2 prefetch distance of the weight tiles on the weight-streaming convolution shapes (isolated)
M="python tools/gemm_micro.py --bias --iters 20"
for pf in 0 8 24 48; do
  TML_B_PREFETCH=$pf $M --B 16 --H 8 --W 8 --Cin 1280 --N 1280 --tag "8^2 1280->1280 pf=$pf"
  TML_B_PREFETCH=$pf $M --B 16 --H 16 --W 16 --Cin 1280 --N 1280 --tag "16^2 1280->1280 pf=$pf"
  TML_B_PREFETCH=$pf $M --B 16 --H 8 --W 8 --Cin 2560 --N 1280 --tag "8^2 2560->1280 pf=$pf"
done
