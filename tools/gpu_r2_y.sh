#!/bin/bash
# round 2, call Y: VAE attention backward with the transposed-operand GEMMs (a_trans on CTA pairs) vs transposed copies
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_decoder.py tests/test_gpu_kernels.py tests/test_gpu_unet.py -x -q > gpurun_out/y_test.log 2>&1; tail -2 gpurun_out/y_test.log
for v in 0 1; do
  TML_NO_ATRANS=$v timeout 300 python bench.py --quick --steps 10 --warmup 3 --gemm_table > gpurun_out/y_512_noatrans$v.json 2> gpurun_out/y_512_noatrans$v.err
  TML_NO_ATRANS=$v timeout 300 python bench.py --quick --res 1024 --batch 16 --steps 5 --warmup 3 --gemm_table > gpurun_out/y_1024_noatrans$v.json 2> gpurun_out/y_1024_noatrans$v.err
done
for f in gpurun_out/y_*.json; do echo $f; cut -c1-170 $f; done
grep -h "attn.dK\|attn.dV" gpurun_out/y_*.err
