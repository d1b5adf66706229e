set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_pmvs_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_gpu_call26.log 2>&1; tail -n 3 gpurun_out/r2/pytest_gpu_call26.log
for vb in 16 32; do for wl in temple47_mu5 temple47_mu7; do
  MVS_K2_VB=$vb python profiles/r2_probe.py --workload $wl --reps 5 --no-probe > gpurun_out/r2/k2_vb${vb}_$wl.log 2>&1; tail -n 2 gpurun_out/r2/k2_vb${vb}_$wl.log
done; done
