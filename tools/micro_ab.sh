M="python tools/gemm_micro.py --B 16 --H 128 --W 128 --Cin 512 --N 512 --bias"
timeout 60 $M --tag pair_full
TML_DBG_NO_EPI=1 timeout 60 $M --tag pair_noepi
TML_DBG_MMA_ONLY=1 timeout 60 $M --tag pair_noloads
TML_NO_SWAP_PAIR=1 timeout 60 $M --tag old_full
timeout 60 $M --resid --gn 1 --tag pair_res_stats
timeout 60 $M --gn 2 --tag pair_gnbwd
