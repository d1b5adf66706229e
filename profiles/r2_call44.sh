set -x
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29519"
run() { # workload, name, env...
wl=$1; name=$2; shift; shift
env "$@" $TR --nproc-per-node 8 bench.py --gpus 8 --steps 60 --workload $wl > gpurun_out/r2/bench_${wl}_n8_$name.json 2> gpurun_out/r2/bench_${wl}_n8_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2/bench_${wl}_n8_$name.json').read().strip().splitlines()[-1])
    print('$wl $name', round(d['ms_per_step'],4), round(d['value']/1e9,3), round(d['roofline']['kernel_ms'],4), d['config']['exchange_verified'], d['config']['rounds_verified'], d['e2e']['matches_device_path'], d['config']['launch'][:20])
except Exception as e: print('$wl $name failed', e)
PY
tail -n 2 gpurun_out/r2/bench_${wl}_n8_$name.err | cut -c1-300
}
run dino48 p1 BENCH_XPARTS=1
run dino48 p2 BENCH_XPARTS=2
run dino48 p4 BENCH_XPARTS=4
BEST=$(python - <<PY
import json
best=(1e9,1)
for p in (1,2,4):
    try:
        d=json.loads(open('gpurun_out/r2/bench_dino48_n8_p%d.json'%p).read().strip().splitlines()[-1])
        best=min(best,(d['ms_per_step'],p))
    except Exception: pass
print(best[1])
PY
)
echo BEST=$BEST
run ring128_1080p p$BEST BENCH_XPARTS=$BEST
