set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_pmvs_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_gpu_call28.log 2>&1; tail -n 3 gpurun_out/r2/pytest_gpu_call28.log
MVS_K2_VB=16 python -m pytest tests/test_pmvs_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_gpu_call28_vb16.log 2>&1; tail -n 3 gpurun_out/r2/pytest_gpu_call28_vb16.log
for vb in 16 32; do for wl in temple47_mu5 temple47_mu7; do
  MVS_K2_VB=$vb python profiles/r2_probe.py --workload $wl --reps 5 --no-probe > gpurun_out/r2/k2_vb${vb}q_$wl.log 2>&1; tail -n 1 gpurun_out/r2/k2_vb${vb}q_$wl.log
done; done
