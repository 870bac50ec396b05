# Round-2 call L: final single-GPU record -- default bench line (as the driver runs it), reference arm, per-GPU batch sweep
# (what strong scaling looks like from one GPU), SDXL configs[2], image-loss mode, universal mode
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
timeout 900 python bench.py --gemm_table > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-400 $O/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/ref.json 2> $O/ref.err; echo "ref rc=$?"; cut -c1-200 $O/ref.json
for b in 8 16 32; do timeout 300 python bench.py --quick --scaling weak --batch $b --steps 10 > $O/b$b.json 2>/dev/null; cat $O/b$b.json; done
timeout 900 python bench.py --res 1024 --batch 16 --steps 5 --warmup 3 --no_cpu_baseline --gemm_table > $O/sdxl.json 2> $O/sdxl.err; echo "sdxl rc=$?"; cut -c1-300 $O/sdxl.json
timeout 600 python bench.py --quick --loss images --micro_batch 8 --steps 3 --warmup 2 > $O/images.json 2> $O/images.err; echo "images rc=$?"; cat $O/images.json
timeout 600 python bench.py --mode universal --dataset 512 --steps 3 --warmup 1 > $O/univ1.json 2> $O/univ1.err; echo "univ rc=$?"; cut -c1-300 $O/univ1.json
