# Round-2 call B: the full GPU test suite (no -x so every failure is listed)
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
timeout 2400 python -m pytest tests -m gpu -q --durations=15 > $O/tests.log 2>&1; echo "tests rc=$?"; tail -40 $O/tests.log
