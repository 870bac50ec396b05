# Round-2 call C: lean epilogue -- kernel suite, encoder parity, thin-GEMM micro timings, bench with table
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/tests.log
M="python tools/gemm_micro.py --B 16 --H 256 --W 256 --Cin 128 --N 256 --k1 --bias"
timeout 60 $M --tag shortcut_1x1_lean
TML_NO_TMA_STORE=1 timeout 60 $M --tag shortcut_1x1_regstore
timeout 60 python tools/gemm_micro.py --B 16 --H 256 --W 256 --Cin 256 --N 128 --k1 --resid --tag shortcut_dgrad_lean
timeout 60 python tools/gemm_micro.py --B 16 --H 128 --W 128 --Cin 256 --N 512 --k1 --bias --tag shortcut_256_512
timeout 60 python tools/gemm_micro.py --B 16 --H 64 --W 64 --Cin 512 --N 1536 --k1 --bias --tag qkv
timeout 300 python bench.py --quick --steps 5 --warmup 3 --gemm_table > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; head -60 $O/bench.err
