# Round-2 call F: attention with the softmax in GEMM epilogues -- parity, then timings (512^2 and 1024^2)
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?"; tail -25 $O/tests.log
timeout 300 python bench.py --quick --steps 5 --warmup 3 --gemm_table > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; grep -E "attn|gemm " $O/bench.err
timeout 600 python bench.py --quick --res 1024 --batch 16 --steps 3 --warmup 2 --gemm_table > $O/sdxl.json 2> $O/sdxl.err; echo "sdxl rc=$?"; cat $O/sdxl.json; grep -E "attn|gemm " $O/sdxl.err
