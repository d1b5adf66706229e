set -x
mkdir -p gpurun_out/r2
timeout 70 python bench.py --workload dino_rounds --steps 5 --warmup 2 > gpurun_out/r2/bench_dino_rounds_final.json 2> gpurun_out/r2/bench_dino_rounds_final.err; tail -c 700 gpurun_out/r2/bench_dino_rounds_final.json; tail -n 2 gpurun_out/r2/bench_dino_rounds_final.err
