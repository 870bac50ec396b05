# Round-2 call U (4 GPUs): strong scaling at N = 4 (16 images per GPU) to complete the 1 / 2 / 4 / 8 table; smoke()
mkdir -p gpurun_out/r2u
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $T bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2u/strong4.json 2> gpurun_out/r2u/strong4.err; echo "strong4 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2u/strong4.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','scaling','clocks')}, d['e2e']['value'], d.get('weak'))
PY
