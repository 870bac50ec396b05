CMD="python bench.py --quick --steps 1 --warmup 1 --batch 16 --micro_batch 16"
timeout 300 $CMD > gpurun_out/p_plain.log 2>&1; echo plain rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1_v2.csv $CMD > gpurun_out/p_ncu1.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:swapped_kernel<1, false, false, false>' -s 2 -c 1 -o gpurun_out/prof_swap_c128_r1 $CMD > gpurun_out/p_ncu2.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:swapped_kernel<4, true, false, true>' -s 2 -c 1 -o gpurun_out/prof_swpair_c512_r1 $CMD > gpurun_out/p_ncu3.log 2>&1; echo rc=$?
ls -la gpurun_out | grep -i "prof_sw\|launches_r1_v2"
tail -3 gpurun_out/p_ncu2.log
