set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_partitioned_gpu.py tests/test_rounds_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_gpu_call38.log 2>&1; tail -n 25 gpurun_out/r2/pytest_gpu_call38.log
for xm in 1 2; do
MVS_XMODE=$xm BENCH_XPARTS=2 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/r2/bench_dino48_n1_xp2_xm$xm.json 2> gpurun_out/r2/bench_dino48_n1_xp2_xm$xm.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2/bench_dino48_n1_xp2_xm$xm.json').read().strip().splitlines()[-1])
    print('xm$xm', d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['exchange_verified'], d['e2e']['matches_device_path'])
except Exception as e: print('xm$xm failed', e)
PY
tail -n 3 gpurun_out/r2/bench_dino48_n1_xp2_xm$xm.err
done
