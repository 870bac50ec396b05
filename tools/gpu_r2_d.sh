# Round-2 call D: where are the 128-channel and 64^2 layers bound?  (operand-load ablations, same box) + lean A/B on the whole step
M="python tools/gemm_micro.py --B 16 --H 512 --W 512 --Cin 128 --N 128 --bias"
timeout 60 $M --gn 1 --tag sw128_full
TML_DBG_MMA_ONLY=2 timeout 60 $M --gn 1 --tag sw128_no_weight_loads
TML_DBG_MMA_ONLY=3 timeout 60 $M --gn 1 --tag sw128_no_row_loads
TML_DBG_MMA_ONLY=1 timeout 60 $M --gn 1 --tag sw128_no_loads
TML_DBG_NO_EPI=1 timeout 60 $M --gn 1 --tag sw128_no_epilogue
timeout 60 $M --gn 2 --tag sw128_gnbwd_full
TML_DBG_MMA_ONLY=1 timeout 60 $M --gn 2 --tag sw128_gnbwd_no_loads
TML_DBG_NO_EPI=1 timeout 60 $M --gn 2 --tag sw128_gnbwd_no_epilogue
M="python tools/gemm_micro.py --B 32 --H 64 --W 64 --Cin 512 --N 512 --bias"
timeout 60 $M --gn 1 --tag pm64_full
TML_DBG_MMA_ONLY=1 timeout 60 $M --gn 1 --tag pm64_no_loads
TML_DBG_NO_EPI=1 timeout 60 $M --gn 1 --tag pm64_no_epilogue
TML_PAIR=0 timeout 60 $M --gn 1 --tag pm64_nopair
mkdir -p gpurun_out/r2d
timeout 300 python bench.py --quick --steps 5 --warmup 3 > gpurun_out/r2d/lean.json 2>/dev/null; cat gpurun_out/r2d/lean.json
TML_NO_TMA_STORE=1 timeout 300 python bench.py --quick --steps 5 --warmup 3 > gpurun_out/r2d/reg.json 2>/dev/null; cat gpurun_out/r2d/reg.json
timeout 300 python bench.py --quick --steps 5 --warmup 3 > gpurun_out/r2d/lean2.json 2>/dev/null; cat gpurun_out/r2d/lean2.json
