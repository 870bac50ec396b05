# Round-2 call S: fused input GN with eight transform warps -- correctness of the kernel suite, isolated timing, whole step A/B
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm_suite" > $O/ktests.log 2>&1; echo "kernel tests rc=$?"; tail -3 $O/ktests.log
M="python tools/gemm_micro.py --B 16 --H 256 --W 256 --Cin 256 --N 256 --bias --gn 1"
timeout 60 $M --tag swpair256_plain
timeout 60 $M --xf --tag swpair256_xf
M="python tools/gemm_micro.py --B 16 --H 128 --W 128 --Cin 512 --N 512 --bias --gn 1"
timeout 60 $M --tag swpair512_plain
timeout 60 $M --xf --tag swpair512_xf
B="python bench.py --quick --steps 5 --warmup 3"
timeout 300 $B 2>/dev/null | cut -c1-150
TML_NO_FUSE_INGN=1 timeout 300 $B 2>/dev/null | cut -c1-150
timeout 300 $B 2>/dev/null | cut -c1-150
TML_NO_FUSE_INGN=1 timeout 300 $B 2>/dev/null | cut -c1-150
