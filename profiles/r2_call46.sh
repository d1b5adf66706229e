set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu_n1_final2.log 2>&1; tail -n 3 gpurun_out/r2/pytest_gpu_n1_final2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke_final2.log 2>&1; tail -n 2 gpurun_out/r2/smoke_final2.log
