# Round-2 call P: the PyTorch library path on the same B200 (informational), clock-sampler check on a short region, default bench
mkdir -p gpurun_out/r2p
TML_LIBRARY_BASELINE=1 timeout 600 python -m pytest tests/test_gpu_library_baseline.py -m gpu -s -q 2>&1 | tail -6
cp gpurun_out/library_baseline.json gpurun_out/r2p/ 2>/dev/null
TML_LIBRARY_BASELINE=1 TML_LIBRARY_BATCH=32 timeout 600 python -m pytest tests/test_gpu_library_baseline.py -m gpu -s -q 2>&1 | grep '^{'
timeout 300 python bench.py --scaling weak --batch 8 --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/r2p/short.json 2> gpurun_out/r2p/short.err; echo "short rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2p/short.json').read().strip().split('\n')[-1]); print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'])"
timeout 600 python bench.py > gpurun_out/r2p/bench.json 2> gpurun_out/r2p/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2p/bench.json').read().strip().split('\n')[-1]); print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'], d['roofline']['frac'])"
