set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_partitioned_gpu.py tests/test_rounds_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_gpu_call39.log 2>&1; tail -n 25 gpurun_out/r2/pytest_gpu_call39.log
for xp in 2 4; do
BENCH_XPARTS=$xp python bench.py --no-cpu-baseline --steps 100 > gpurun_out/r2/bench_dino48_n1_gate$xp.json 2> gpurun_out/r2/bench_dino48_n1_gate$xp.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2/bench_dino48_n1_gate$xp.json').read().strip().splitlines()[-1])
    print('gate$xp', d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['exchange_verified'], d['e2e']['matches_device_path'], d['config']['launch'])
except Exception as e: print('gate$xp failed', e)
PY
tail -n 3 gpurun_out/r2/bench_dino48_n1_gate$xp.err
done
