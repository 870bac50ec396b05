# Round-2 call H: tanh-form silu' in the fused GroupNorm-backward epilogue + residual prefetch in the lean epilogue: parity, then timings
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?"; tail -8 $O/tests.log
M="python tools/gemm_micro.py --B 16 --H 512 --W 512 --Cin 128 --N 128"
timeout 60 $M --gn 2 --tag "sw128 gnbwd tanh"
timeout 60 python tools/gemm_micro.py --B 16 --H 256 --W 256 --Cin 256 --N 256 --gn 2 --tag "swpair256 gnbwd tanh"
timeout 60 python tools/gemm_micro.py --B 16 --H 256 --W 256 --Cin 256 --N 128 --k1 --resid --tag shortcut_dgrad_lean_prefetch
timeout 300 python bench.py --quick --steps 5 --warmup 3 --gemm_table > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; head -12 $O/bench.err; grep -E "attn.dS|attn.qk|shortcut" $O/bench.err
