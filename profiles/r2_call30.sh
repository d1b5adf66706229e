set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_pmvs_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_gpu_call30.log 2>&1; tail -n 3 gpurun_out/r2/pytest_gpu_call30.log
for wl in temple47_mu5 temple47_mu7; do
  python profiles/r2_probe.py --workload $wl --reps 5 --no-probe > gpurun_out/r2/k2_nounroll_$wl.log 2>&1; tail -n 1 gpurun_out/r2/k2_nounroll_$wl.log
done
