set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu_call5.log 2>&1; tail -8 gpurun_out/r2/pytest_gpu_call5.log
python profiles/r2_rounds_trace.py > gpurun_out/r2/rounds_trace.log 2>&1; tail -2 gpurun_out/r2/rounds_trace.log
python bench.py --workload dino_rounds --steps 10 --warmup 2 > gpurun_out/r2/bench_dino_rounds.json 2> gpurun_out/r2/bench_dino_rounds.err; tail -c 1200 gpurun_out/r2/bench_dino_rounds.json; tail -5 gpurun_out/r2/bench_dino_rounds.err
for mb in 2 3; do for wl in ring128_1080p ring256_4k; do MVS_K1_MINB=$mb python profiles/r2_probe.py --workload $wl > gpurun_out/r2/probe_${wl}_minb$mb.json 2>&1; tail -1 gpurun_out/r2/probe_${wl}_minb$mb.json; done; done
python bench.py --workload ring128_1080p --no-cpu-baseline > gpurun_out/r2/bench_ring128_n1.json 2> gpurun_out/r2/bench_ring128_n1.err; tail -c 800 gpurun_out/r2/bench_ring128_n1.json; tail -3 gpurun_out/r2/bench_ring128_n1.err
