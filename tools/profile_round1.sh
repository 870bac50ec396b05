# Round-1 ncu evidence (run on the GPU box through gpurun; each ncu pass follows a plain run of the same command).
CMD="python bench.py --quick --steps 1 --warmup 1 --batch 16 --micro_batch 16"
timeout 300 $CMD > gpurun_out/p_plain.log 2>&1; echo plain rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1_v4.csv $CMD > gpurun_out/p_ncu1.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:swapped_kernel<.int.1, .bool.0, .bool.0, .bool.0>' -s 2 -c 1 -o gpurun_out/prof_swap_c128_r1 $CMD > gpurun_out/p_ncu2.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:swapped_kernel<.int.4, .bool.1, .bool.0, .bool.1>' -s 2 -c 1 -o gpurun_out/prof_swpair_c512_r1 $CMD > gpurun_out/p_ncu3.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_bwd_apply -s 22 -c 1 -o gpurun_out/prof_gnbwd_apply_r1 $CMD > gpurun_out/p_ncu4.log 2>&1; echo rc=$?
ls -la gpurun_out | grep -i "prof_"
