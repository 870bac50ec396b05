M="python tools/gemm_micro.py --B 16 --H 512 --W 512 --Cin 128 --N 128 --bias"
timeout 60 $M --gn 1 --tag sw128_full
TML_DBG_NO_EPI=2 timeout 60 $M --gn 1 --tag sw128_nostore
TML_DBG_NO_EPI=1 timeout 60 $M --gn 1 --tag sw128_noepi
TML_DBG_MMA_ONLY=1 timeout 60 $M --gn 1 --tag sw128_noloads
TML_DBG_MMA_ONLY=1 TML_DBG_NO_EPI=1 timeout 60 $M --gn 1 --tag sw128_mmaonly
M="python tools/gemm_micro.py --B 16 --H 256 --W 256 --Cin 256 --N 256 --bias"
timeout 60 $M --gn 1 --tag pair256_full
TML_DBG_NO_EPI=2 timeout 60 $M --gn 1 --tag pair256_nostore
TML_DBG_NO_EPI=1 timeout 60 $M --gn 1 --tag pair256_noepi
