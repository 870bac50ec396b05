# Round-2 call M (8 GPUs): strong scaling of the fixed 64-image batch (+ weak in the same run) and the universal mode of configs[3]
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $T bench.py --gpus 8 --steps 10 --warmup 3 > $O/strong8.json 2> $O/strong8.err; echo "strong8 rc=$?"; grep '^{' $O/strong8.json | cut -c1-300; tail -2 $O/strong8.err
NCCL_DEBUG=INFO timeout 600 $T bench.py --gpus 8 --mode universal --dataset 512 --steps 5 --warmup 2 > $O/univ8.json 2> $O/univ8.err; echo "univ8 rc=$?"; grep '^{' $O/univ8.json; grep -i -m3 "nvls\|Using network\|channels" $O/univ8.json $O/univ8.err | cut -c1-200
