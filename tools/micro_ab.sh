M="python tools/gemm_micro.py --B 16 --H 256 --W 256 --Cin 128 --N 256 --k1 --bias"
timeout 60 $M --tag sc_bias
M="python tools/gemm_micro.py --B 16 --H 64 --W 64 --Cin 512 --N 512 --bias"
timeout 60 $M --gn 1 --tag c512_64_stats
timeout 60 $M --gn 1 --resid --tag c512_64_res_stats
timeout 60 $M --gn 2 --tag c512_64_gnbwd
TML_DBG_NO_EPI=1 timeout 60 $M --gn 1 --tag c512_64_noepi
