# Round-2 call J: pitch-66 halo mode for the 64^2 stage -- kernel suite, micro timings, parity, bench A/B
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm_suite" > $O/ktests.log 2>&1; echo "kernel tests rc=$?"; grep -E "h66|FAIL|passed|failed" $O/ktests.log | head -20
M="python tools/gemm_micro.py --B 32 --H 64 --W 64 --Cin 512 --N 512 --bias"
timeout 60 $M --gn 1 --tag pm64_h66
TML_NO_H66=1 timeout 60 $M --gn 1 --tag pm64_tap_by_tap
timeout 60 $M --gn 2 --tag pm64_h66_gnbwd
TML_NO_H66=1 timeout 60 $M --gn 2 --tag pm64_tap_by_tap_gnbwd
TML_DBG_MMA_ONLY=1 timeout 60 $M --gn 1 --tag pm64_h66_no_loads
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?"; tail -6 $O/tests.log
B="python bench.py --quick --steps 5 --warmup 3"
timeout 300 $B 2>/dev/null | cut -c1-150
TML_NO_H66=1 timeout 300 $B 2>/dev/null | cut -c1-150
timeout 300 $B --gemm_table 2> $O/bench.err | cut -c1-150; grep -E " 512  4608 " $O/bench.err
