set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu_n1.log 2>&1; tail -15 gpurun_out/r2/pytest_gpu_n1.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke.log 2>&1; tail -3 gpurun_out/r2/smoke.log
python bench.py --workload dino_rounds --steps 10 --warmup 2 > gpurun_out/r2/bench_dino_rounds.json 2> gpurun_out/r2/bench_dino_rounds.err; tail -c 3000 gpurun_out/r2/bench_dino_rounds.json; tail -5 gpurun_out/r2/bench_dino_rounds.err
python bench.py > gpurun_out/r2/bench_dino48.json 2> gpurun_out/r2/bench_dino48.err; tail -c 3000 gpurun_out/r2/bench_dino48.json; tail -5 gpurun_out/r2/bench_dino48.err
for mb in 2 4; do MVS_K6_MINB=$mb python profiles/r2_probe.py --workload dino48 > gpurun_out/r2/probe_dino48_minb$mb.json 2>&1; tail -1 gpurun_out/r2/probe_dino48_minb$mb.json; done
