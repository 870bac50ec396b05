#!/bin/bash
# round 2 (third session), calls C (before) and H (after the loads-in-flight changes): DRAM throughput of every non-GEMM kernel of one native-UNet diffusion step
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none --csv -k regex:'ln_|gng_|geglu|head_|copy_cols|gn_|add_bf16|upsample|transpose|prep' \
  --log-file gpurun_out/r3h_elem.csv python bench.py --mode diffusion --batch 8 --unet native --steps 1 --warmup 0 --no_cpu_baseline > gpurun_out/r3h_ncu.log 2>&1
tail -3 gpurun_out/r3h_ncu.log; wc -l gpurun_out/r3h_elem.csv
