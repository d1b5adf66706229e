set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu_final.log 2>&1; tail -4 gpurun_out/r2/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke.log 2>&1; tail -2 gpurun_out/r2/smoke.log
for wl in temple47_mu5 temple47_mu7; do python bench.py --workload $wl --steps 20 > gpurun_out/r2/bench_${wl}_n1.json 2> gpurun_out/r2/bench_${wl}_n1.err; tail -c 300 gpurun_out/r2/bench_${wl}_n1.json; tail -2 gpurun_out/r2/bench_${wl}_n1.err; done
python bench.py > gpurun_out/r2/bench_dino48_n1_final.json 2> gpurun_out/r2/bench_dino48_n1_final.err; tail -c 300 gpurun_out/r2/bench_dino48_n1_final.json
