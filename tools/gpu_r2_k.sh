# Round-2 call K: full GPU suite after the pitch-66 halo mode + final ncu evidence
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
timeout 1500 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -6 $O/tests.log
CMD="python bench.py --quick --steps 1 --warmup 1 --batch 16 --micro_batch 16 --scaling weak"
timeout 300 $CMD > $O/p_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches.csv $CMD > $O/p_ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_gemm_tcgen05_kernel<.bool.1, .int.3>' -s 1 -c 1 -o $O/prof_attn_qk_exp $CMD > $O/p_ncu2.log 2>&1; echo rc=$?
M="python tools/gemm_micro.py --B 16 --H 256 --W 256 --Cin 128 --N 256 --k1 --bias --iters 3"
timeout 120 $M > $O/m_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tcgen05 -s 3 -c 1 -o $O/prof_thin_shortcut_lean $M > $O/p_ncu3.log 2>&1; echo rc=$?
M="python tools/gemm_micro.py --B 32 --H 64 --W 64 --Cin 512 --N 512 --bias --gn 1 --iters 3"
timeout 120 $M > $O/m2_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tcgen05 -s 3 -c 1 -o $O/prof_h66_c512_64 $M > $O/p_ncu4.log 2>&1; echo rc=$?
ls -la $O | head -30
