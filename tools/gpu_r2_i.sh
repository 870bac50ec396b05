# Round-2 call I: after splitting the attention modes out of the plain lean epilogue: parity + timings
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?"; tail -8 $O/tests.log
timeout 60 python tools/gemm_micro.py --B 16 --H 256 --W 256 --Cin 128 --N 256 --k1 --bias --tag shortcut_1x1_lean
timeout 60 python tools/gemm_micro.py --B 16 --H 256 --W 256 --Cin 256 --N 128 --k1 --resid --tag shortcut_dgrad_lean
timeout 300 python bench.py --quick --steps 5 --warmup 3 --gemm_table > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; grep -E "attn.dS|attn.qk|shortcut|conv_in.dgrad" $O/bench.err
timeout 600 python bench.py --quick --res 1024 --batch 16 --steps 3 --warmup 2 > $O/sdxl.json 2> $O/sdxl.err; cat $O/sdxl.json
