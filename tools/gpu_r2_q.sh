# Round-2 call Q: GroupNorm + SiLU of the input fused into the CTA-pair 3x3 kernel's operand path -- kernel suite, parity, A/B
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm_suite" > $O/ktests.log 2>&1; echo "kernel tests rc=$?"; grep -E "xf|FAIL|passed|failed|hang|rror" $O/ktests.log | head -20
timeout 1500 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?"; tail -6 $O/tests.log
B="python bench.py --quick --steps 5 --warmup 3"
timeout 300 $B 2>/dev/null | cut -c1-150
TML_NO_FUSE_INGN=1 timeout 300 $B 2>/dev/null | cut -c1-150
timeout 300 $B --gemm_table 2> $O/bench.err | cut -c1-150; grep -E " 31[0-9][0-9] " $O/bench.err
TML_NO_FUSE_INGN=1 timeout 300 $B --gemm_table 2> $O/bench_off.err | cut -c1-150; grep -E " 30[0-9][0-9] " $O/bench_off.err
