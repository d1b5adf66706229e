set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu_final.log 2>&1; tail -n 4 gpurun_out/r2/pytest_gpu_final.log
python bench.py > gpurun_out/r2/bench_dino48_n1_final.json 2> gpurun_out/r2/bench_dino48_n1_final.err; tail -c 300 gpurun_out/r2/bench_dino48_n1_final.json
python bench.py --workload dino_rounds --steps 10 --warmup 2 > gpurun_out/r2/bench_dino_rounds.json 2> gpurun_out/r2/bench_dino_rounds.err; tail -c 300 gpurun_out/r2/bench_dino_rounds.json
python bench.py --workload ring128_1080p --no-cpu-baseline --steps 50 > gpurun_out/r2/bench_ring128_n1_final.json 2> gpurun_out/r2/bench_ring128_n1_final.err; tail -c 300 gpurun_out/r2/bench_ring128_n1_final.json
