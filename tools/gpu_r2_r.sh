# Round-2 call R: why is the fused-input-GN pair kernel slow?  isolated timing + ncu source view
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
M="python tools/gemm_micro.py --B 16 --H 256 --W 256 --Cin 256 --N 256 --bias --gn 1"
timeout 60 $M --tag swpair256_plain
timeout 60 $M --xf --tag swpair256_xf
M="$M --xf --iters 3"
timeout 120 $M > $O/m_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_swapped -s 3 -c 1 -o $O/prof_xf_c256 $M > $O/p_ncu.log 2>&1; echo rc=$?
