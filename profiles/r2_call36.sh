set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu_n1_final.log 2>&1; tail -n 3 gpurun_out/r2/pytest_gpu_n1_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke_final.log 2>&1; tail -n 2 gpurun_out/r2/smoke_final.log
python bench.py > gpurun_out/r2/bench_default_final.json 2> gpurun_out/r2/bench_default_final.err; tail -c 600 gpurun_out/r2/bench_default_final.json; tail -n 2 gpurun_out/r2/bench_default_final.err
python bench.py --impl reference > gpurun_out/r2/bench_reference_final.json 2> gpurun_out/r2/bench_reference_final.err; tail -c 400 gpurun_out/r2/bench_reference_final.json
