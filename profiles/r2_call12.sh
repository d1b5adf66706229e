set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_score_gpu.py tests/test_dinoring_full_gpu.py tests/test_rounds_gpu.py tests/test_compact_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_gpu_call12.log 2>&1; tail -12 gpurun_out/r2/pytest_gpu_call12.log
for wl in ring128_1080p ring256_4k; do
  python profiles/r2_probe.py --workload $wl > gpurun_out/r2/probe_${wl}_k8.json 2>&1; tail -1 gpurun_out/r2/probe_${wl}_k8.json
  MVS_K8=0 python profiles/r2_probe.py --workload $wl > gpurun_out/r2/probe_${wl}_k8off.json 2>&1; tail -1 gpurun_out/r2/probe_${wl}_k8off.json
  MVS_K8_MINB=3 python profiles/r2_probe.py --workload $wl > gpurun_out/r2/probe_${wl}_k8_minb3.json 2>&1; tail -1 gpurun_out/r2/probe_${wl}_k8_minb3.json
  MVS_K8_PER=4 python profiles/r2_probe.py --workload $wl > gpurun_out/r2/probe_${wl}_k8_per4.json 2>&1; tail -1 gpurun_out/r2/probe_${wl}_k8_per4.json
  MVS_K8_PER=16 python profiles/r2_probe.py --workload $wl > gpurun_out/r2/probe_${wl}_k8_per16.json 2>&1; tail -1 gpurun_out/r2/probe_${wl}_k8_per16.json
done
