# A/B timings of single convolution shapes in isolation (run on the GPU box: gpurun -- 'bash tools/micro_ab.sh').
# Every setting is its own process because the kernels read their environment switches once.
M="python tools/gemm_micro.py --B 16 --H 128 --W 128 --Cin 512 --N 512 --bias"
timeout 60 $M --tag pair_full                                   # operand-swapped kernel, CTA pairs
TML_DBG_NO_EPI=1 timeout 60 $M --tag pair_noepi                 # ... without the epilogue
TML_DBG_MMA_ONLY=1 timeout 60 $M --tag pair_noloads             # ... without operand loads after the first ring pass
TML_NO_SWAP_PAIR=1 timeout 60 $M --tag pixel_major_full         # the pixel-major kernel on the same shape
timeout 60 $M --resid --gn 1 --tag pair_res_stats
timeout 60 $M --gn 2 --tag pair_gnbwd
M="python tools/gemm_micro.py --B 16 --H 512 --W 512 --Cin 128 --N 128 --bias"
timeout 60 $M --gn 1 --tag sw128_full                           # single-CTA swapped kernel, 128 channels
TML_DBG_NO_EPI=2 timeout 60 $M --gn 1 --tag sw128_nostore
TML_DBG_NO_EPI=1 timeout 60 $M --gn 1 --tag sw128_noepi
TML_NO_SWAP=1 timeout 60 $M --gn 1 --tag pixel_major_128
M="python tools/gemm_micro.py --B 16 --H 256 --W 256 --Cin 128 --N 256 --k1 --bias"
timeout 60 $M --tag shortcut_1x1                                # a thin GEMM: bound by the epilogue's store path
TML_DBG_NO_EPI=1 timeout 60 $M --tag shortcut_1x1_noepi
