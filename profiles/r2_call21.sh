set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_score_gpu.py tests/test_dinoring_full_gpu.py tests/test_rounds_gpu.py tests/test_shim_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_gpu_call21.log 2>&1; tail -n 4 gpurun_out/r2/pytest_gpu_call21.log
for wl in ring128_1080p ring256_4k; do
  python profiles/r2_probe.py --workload $wl > gpurun_out/r2/probe_${wl}_even.json 2>&1; tail -n 1 gpurun_out/r2/probe_${wl}_even.json
  MVS_K1_EVEN_BINS=0 python profiles/r2_probe.py --workload $wl > gpurun_out/r2/probe_${wl}_noeven.json 2>&1; tail -n 1 gpurun_out/r2/probe_${wl}_noeven.json
done
