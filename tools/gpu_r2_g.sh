# Round-2 call G (2 GPUs): strong-scaling line (+ weak in the same run), universal mode with the NCCL all-reduce, reference arm under torchrun
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 2 --steps 5 --warmup 3 > $O/strong2.json 2> $O/strong2.err; echo "strong2 rc=$?"; cat $O/strong2.json | cut -c1-1500; tail -3 $O/strong2.err
timeout 600 $T bench.py --gpus 2 --mode universal --dataset 128 --steps 3 --warmup 1 > $O/univ2.json 2> $O/univ2.err; echo "univ2 rc=$?"; cat $O/univ2.json; tail -3 $O/univ2.err
timeout 600 $T -m tml_image_editing_defense_b200.main --universal --num_images 16 --max_train_steps 3 --train_batch_size 8 --output_dir $O/univ_cli > $O/univ_cli.log 2>&1; echo "cli rc=$?"; grep -h "mode" $O/univ_cli.log
timeout 300 $T bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > $O/ref2.json 2> $O/ref2.err; echo "ref2 rc=$?"; cat $O/ref2.json | cut -c1-400
