# Round-2 call N: transposed copies of P~ / dS written by the attention epilogues (no big transpose kernels) -- parity + timings
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
timeout 1500 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?"; tail -8 $O/tests.log
timeout 300 python bench.py --quick --steps 5 --warmup 3 --gemm_table > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; grep -E "attn" $O/bench.err
timeout 600 python bench.py --quick --res 1024 --batch 16 --steps 3 --warmup 2 --gemm_table > $O/sdxl.json 2> $O/sdxl.err; cat $O/sdxl.json; grep -E "attn" $O/sdxl.err
