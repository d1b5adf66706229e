set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_pmvs_gpu.py tests/test_torch_ops.py -m gpu -x -q > gpurun_out/r2/pytest_gpu_call18.log 2>&1; tail -n 8 gpurun_out/r2/pytest_gpu_call18.log
for wl in temple47_mu5 temple47_mu7; do python bench.py --workload $wl --steps 20 > gpurun_out/r2/bench_${wl}_n1.json 2> gpurun_out/r2/bench_${wl}_n1.err; tail -c 300 gpurun_out/r2/bench_${wl}_n1.json; tail -n 2 gpurun_out/r2/bench_${wl}_n1.err; done
