set -x
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29519"
MVS_XMODE=1 BENCH_XPARTS=2 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 60 > gpurun_out/r2/bench_dino48_n8_p2_prio.json 2> gpurun_out/r2/bench_dino48_n8_p2_prio.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2/bench_dino48_n8_p2_prio.json').read().strip().splitlines()[-1])
    print('dino48 p2 prio', round(d['ms_per_step'],4), round(d['value']/1e9,3), round(d['roofline']['kernel_ms'],4), d['config']['exchange_verified'], d['config']['rounds_verified'], d['e2e']['matches_device_path'], d['config']['launch'][:20])
except Exception as e: print('failed', e)
PY
tail -n 2 gpurun_out/r2/bench_dino48_n8_p2_prio.err | cut -c1-300
