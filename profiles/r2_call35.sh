set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_score_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_gpu_call35.log 2>&1; tail -n 3 gpurun_out/r2/pytest_gpu_call35.log
python bench.py --workload ring128_1080p --steps 50 > gpurun_out/r2/bench_ring128_n1.json 2> gpurun_out/r2/bench_ring128_n1.err; tail -c 300 gpurun_out/r2/bench_ring128_n1.json; tail -n 2 gpurun_out/r2/bench_ring128_n1.err
