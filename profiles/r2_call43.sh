set -x
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29517"
python -m pytest tests/test_rounds_multi_gpu.py -m gpu -x -q -k "2" > gpurun_out/r2/pytest_multi_gpu_n2_parts.log 2>&1; tail -n 5 gpurun_out/r2/pytest_multi_gpu_n2_parts.log
run() { # name, env...
name=$1; shift
env "$@" $TR --nproc-per-node 2 bench.py --gpus 2 --steps 100 > gpurun_out/r2/bench_dino48_n2_$name.json 2> gpurun_out/r2/bench_dino48_n2_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2/bench_dino48_n2_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), d['config']['exchange_verified'], d['config']['rounds_verified'], d['e2e']['matches_device_path'], d['config']['launch'][:20])
except Exception as e: print('$name failed', e)
PY
tail -n 2 gpurun_out/r2/bench_dino48_n2_$name.err
}
run p1 BENCH_XPARTS=1
run p2 BENCH_XPARTS=2
run p4 BENCH_XPARTS=4
