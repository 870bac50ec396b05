M="python tools/gemm_micro.py --B 16 --H 512 --W 512 --Cin 64 --N 128 --bias"
timeout 60 $M --tag sw_k576_bias
timeout 60 $M --gn 1 --tag sw_k576_bias_stats
timeout 60 $M --gn 1 --resid --tag sw_k576_res_stats
M="python tools/gemm_micro.py --B 16 --H 512 --W 512 --Cin 128 --N 128 --bias"
timeout 60 $M --gn 1 --tag sw_k1152_bias_stats
timeout 60 $M --gn 1 --resid --tag sw_k1152_res_stats
