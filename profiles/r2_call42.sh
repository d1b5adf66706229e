set -x
mkdir -p gpurun_out/r2
run() { # name, env...
name=$1; shift
env "$@" python bench.py --no-cpu-baseline --steps 100 > gpurun_out/r2/bench_dino48_n1_$name.json 2> gpurun_out/r2/bench_dino48_n1_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2/bench_dino48_n1_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), d['config']['exchange_verified'], d['e2e']['matches_device_path'], d['config']['launch'][:20])
except Exception as e: print('$name failed', e)
PY
tail -n 2 gpurun_out/r2/bench_dino48_n1_$name.err
}
run c16_p1 BENCH_XPARTS=1 MVS_K1_CAPMUL=16
run c4_p1 BENCH_XPARTS=1 MVS_K1_CAPMUL=4
run s_p2 BENCH_XPARTS=2
run s_p2_m1 BENCH_XPARTS=2 MVS_XMODE=1
run s_p4_m1 BENCH_XPARTS=4 MVS_XMODE=1
