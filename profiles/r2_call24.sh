set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_rounds_gpu.py tests/test_shim_gpu.py tests/test_dinoring_full_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_gpu_call24.log 2>&1; tail -n 4 gpurun_out/r2/pytest_gpu_call24.log
python bench.py --workload dino_rounds --steps 10 --warmup 2 > gpurun_out/r2/bench_dino_rounds.json 2> gpurun_out/r2/bench_dino_rounds.err; tail -c 300 gpurun_out/r2/bench_dino_rounds.json; tail -n 3 gpurun_out/r2/bench_dino_rounds.err
python bench.py --no-cpu-baseline > gpurun_out/r2/bench_dino48_n1_final.json 2> gpurun_out/r2/bench_dino48_n1_final.err; tail -c 300 gpurun_out/r2/bench_dino48_n1_final.json
python profiles/r2_rounds_trace.py > gpurun_out/r2/rounds_trace.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2/rounds_launches.csv python profiles/r2_rounds_trace.py > gpurun_out/r2/rounds_trace_ncu.log 2>&1
tail -n 2 gpurun_out/r2/rounds_trace.log
