set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu_call8.log 2>&1; tail -6 gpurun_out/r2/pytest_gpu_call8.log
for wl in ring128_1080p ring256_4k; do python profiles/r2_probe.py --workload $wl > gpurun_out/r2/probe_${wl}_dyn.json 2>&1; tail -1 gpurun_out/r2/probe_${wl}_dyn.json; done
python bench.py --workload dino_rounds --steps 10 --warmup 2 > gpurun_out/r2/bench_dino_rounds.json 2> gpurun_out/r2/bench_dino_rounds.err; tail -c 700 gpurun_out/r2/bench_dino_rounds.json; tail -3 gpurun_out/r2/bench_dino_rounds.err
python bench.py --workload ring128_1080p --no-cpu-baseline > gpurun_out/r2/bench_ring128_n1_dyn.json 2> gpurun_out/r2/bench_ring128_n1_dyn.err; tail -c 500 gpurun_out/r2/bench_ring128_n1_dyn.json
