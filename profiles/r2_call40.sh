set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_partitioned_gpu.py tests/test_rounds_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_gpu_call40.log 2>&1; tail -n 5 gpurun_out/r2/pytest_gpu_call40.log
run() { # name, env...
name=$1; shift
env "$@" python bench.py --no-cpu-baseline --steps 100 > gpurun_out/r2/bench_dino48_n1_$name.json 2> gpurun_out/r2/bench_dino48_n1_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2/bench_dino48_n1_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), d['config']['exchange_verified'], d['e2e']['matches_device_path'])
except Exception as e: print('$name failed', e)
PY
tail -n 2 gpurun_out/r2/bench_dino48_n1_$name.err
}
run x1 BENCH_XPARTS=1
run x2_m3 BENCH_XPARTS=2 MVS_XMODE=3
run x2_g16 BENCH_XPARTS=2 MVS_XGRID=16
run x2_g32 BENCH_XPARTS=2 MVS_XGRID=32
run x2_g64 BENCH_XPARTS=2 MVS_XGRID=64
run x2_g128 BENCH_XPARTS=2 MVS_XGRID=128
run x4_g64 BENCH_XPARTS=4 MVS_XGRID=64
