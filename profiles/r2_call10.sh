set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu_final.log 2>&1; tail -6 gpurun_out/r2/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke.log 2>&1; tail -2 gpurun_out/r2/smoke.log
python bench.py > gpurun_out/r2/bench_dino48_n1_final.json 2> gpurun_out/r2/bench_dino48_n1_final.err; tail -c 2500 gpurun_out/r2/bench_dino48_n1_final.json; tail -3 gpurun_out/r2/bench_dino48_n1_final.err
MVS_K6_MINB=3 python profiles/r2_probe.py --workload dino48 > gpurun_out/r2/probe_dino48_per8_minb3.json 2>&1; tail -1 gpurun_out/r2/probe_dino48_per8_minb3.json
python profiles/r2_probe.py --workload temple47_a > gpurun_out/r2/probe_temple47_a_final.json 2>&1; tail -1 gpurun_out/r2/probe_temple47_a_final.json
python bench.py --workload dino_rounds --steps 10 --warmup 2 > gpurun_out/r2/bench_dino_rounds.json 2> gpurun_out/r2/bench_dino_rounds.err; tail -c 500 gpurun_out/r2/bench_dino_rounds.json; tail -3 gpurun_out/r2/bench_dino_rounds.err
python bench.py --impl reference > gpurun_out/r2/bench_dino48_reference_arm.json 2> gpurun_out/r2/bench_reference.err; tail -c 500 gpurun_out/r2/bench_dino48_reference_arm.json
