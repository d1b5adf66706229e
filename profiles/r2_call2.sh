set -x
mkdir -p gpurun_out/r2
nvidia-smi -L
python -m pytest tests/test_score_gpu.py tests/test_dinoring_full_gpu.py tests/test_rounds_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_score.log 2>&1; tail -5 gpurun_out/r2/pytest_score.log
for wl in dino48 temple47_a ring128_1080p ring256_4k; do
  python profiles/r2_probe.py --workload $wl > gpurun_out/r2/probe_$wl.json 2> gpurun_out/r2/probe_$wl.err; tail -1 gpurun_out/r2/probe_$wl.json
done
MVS_K1_LEGACY=1 python profiles/r2_probe.py --workload dino48 > gpurun_out/r2/probe_dino48_legacy.json 2>&1; tail -1 gpurun_out/r2/probe_dino48_legacy.json
export_ncu() {  # $1 = tag
  ncu -i gpurun_out/r2/ncu_$1.ncu-rep --page raw --csv > gpurun_out/r2/ncu_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2/ncu_$1.ncu-rep --page source --csv > gpurun_out/r2/ncu_$1_source.csv 2>/dev/null
  rm -f gpurun_out/r2/ncu_$1.ncu-rep
}
for wl in dino48 ring128_1080p ring256_4k temple47_mu5 temple47_mu7; do
  python profiles/r2_probe.py --workload $wl --reps 3 --no-probe > gpurun_out/r2/plain_$wl.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:ncc_score -s 2 -c 1 -f -o gpurun_out/r2/ncu_$wl python profiles/r2_probe.py --workload $wl --reps 3 --no-probe > gpurun_out/r2/ncu_$wl.log 2>&1
  tail -2 gpurun_out/r2/ncu_$wl.log
  export_ncu $wl
done
du -sh gpurun_out
