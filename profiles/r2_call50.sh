mkdir -p gpurun_out/r2
timeout 40 python bench.py --steps 50 > gpurun_out/r2/bench_default_tuned.json 2> gpurun_out/r2/bench_default_tuned.err; tail -c 900 gpurun_out/r2/bench_default_tuned.json; tail -n 2 gpurun_out/r2/bench_default_tuned.err
