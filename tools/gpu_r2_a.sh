# Round-2 call A: GPU tests (incl. parity at the benchmarked geometry), bench lines, per-shape table, ncu evidence.
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --gemm_table > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json | cut -c1-600
timeout 300 python bench.py --quick --scaling weak --batch 8 --steps 10 > $O/b8.json 2> $O/b8.err; echo "b8 rc=$?"; cat $O/b8.json
timeout 300 python bench.py --quick --scaling weak --batch 16 --steps 10 > $O/b16.json 2> $O/b16.err; cat $O/b16.json
timeout 300 python bench.py --quick --scaling weak --batch 32 --steps 5 > $O/b32.json 2> $O/b32.err; cat $O/b32.json
timeout 600 python bench.py --res 1024 --batch 16 --steps 3 --warmup 3 --no_cpu_baseline --gemm_table > $O/sdxl.json 2> $O/sdxl.err; echo "sdxl rc=$?"; cat $O/sdxl.json | cut -c1-400
timeout 600 python bench.py --mode universal --dataset 128 --steps 2 --warmup 1 > $O/univ1.json 2> $O/univ1.err; echo "univ rc=$?"; cat $O/univ1.json | cut -c1-600
CMD="python bench.py --quick --steps 1 --warmup 1 --batch 16 --micro_batch 16 --scaling weak"
timeout 300 $CMD > $O/p_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches.csv $CMD > $O/p_ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_apply_kernel -s 22 -c 1 -o $O/prof_gn_apply $CMD > $O/p_ncu2.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:softmax_rows_kernel -s 1 -c 1 -o $O/prof_softmax $CMD > $O/p_ncu3.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transpose_kernel -s 6 -c 1 -o $O/prof_transpose $CMD > $O/p_ncu4.log 2>&1; echo rc=$?
M="python tools/gemm_micro.py --B 16 --H 256 --W 256 --Cin 128 --N 256 --k1 --bias --iters 3"
timeout 120 $M > $O/m_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tcgen05 -s 3 -c 1 -o $O/prof_thin_shortcut $M > $O/p_ncu5.log 2>&1; echo rc=$?
ls -la $O
