# Round-2 call E: input-row latency of the operand-swapped kernel -- L2 prefetch of the next tile's rows, ring depths
L=tml_image_editing_defense_b200/csrc/build
M="python tools/gemm_micro.py --B 16 --H 512 --W 512 --Cin 128 --N 128 --bias"
for g in 1 2; do
timeout 60 $M --gn $g --tag "sw128 r3w7"
TML_SW_PREFETCH=1 timeout 60 $M --gn $g --tag "sw128 r3w7 +pf"
TML_LIB_PATH=$L/libtml_r4w5.so timeout 60 $M --gn $g --tag "sw128 r4w5"
TML_LIB_PATH=$L/libtml_r4w5.so TML_SW_PREFETCH=1 timeout 60 $M --gn $g --tag "sw128 r4w5 +pf"
TML_LIB_PATH=$L/libtml_r4w4.so TML_SW_PREFETCH=1 timeout 60 $M --gn $g --tag "sw128 r4w4 +pf"
done
M="python tools/gemm_micro.py --B 16 --H 256 --W 256 --Cin 256 --N 256 --bias"
timeout 60 $M --gn 1 --tag "swpair256"
TML_SW_PREFETCH=1 timeout 60 $M --gn 1 --tag "swpair256 +pf"
M="python tools/gemm_micro.py --B 16 --H 128 --W 128 --Cin 512 --N 512 --bias"
timeout 60 $M --gn 1 --tag "swpair512"
TML_SW_PREFETCH=1 timeout 60 $M --gn 1 --tag "swpair512 +pf"
mkdir -p gpurun_out/r2e
B="python bench.py --quick --steps 5 --warmup 3"
timeout 300 $B 2>/dev/null | cut -c1-140
TML_SW_PREFETCH=1 timeout 300 $B 2>/dev/null | cut -c1-140
TML_LIB_PATH=$L/libtml_r4w5.so TML_SW_PREFETCH=1 timeout 300 $B 2>/dev/null | cut -c1-140
timeout 300 $B 2>/dev/null | cut -c1-140
TML_SW_PREFETCH=1 timeout 300 $B 2>/dev/null | cut -c1-140
