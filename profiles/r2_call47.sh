set -x
mkdir -p gpurun_out/r2
timeout 140 python bench.py > gpurun_out/r2/bench_default_final2.json 2> gpurun_out/r2/bench_default_final2.err; tail -c 400 gpurun_out/r2/bench_default_final2.json
