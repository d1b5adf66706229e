set -x
mkdir -p gpurun_out/r2
nvidia-smi -L
for wl in dino48 ring128_1080p ring256_4k; do
  python profiles/r2_probe.py --workload $wl > gpurun_out/r2/probe_$wl.json 2> gpurun_out/r2/probe_$wl.err; tail -1 gpurun_out/r2/probe_$wl.json
done
for wl in ring128_1080p ring256_4k temple47_mu5 temple47_mu7; do
  python profiles/r2_probe.py --workload $wl --reps 3 --no-probe > gpurun_out/r2/plain_$wl.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:ncc_score -s 2 -c 1 -f -o gpurun_out/r2/ncu_$wl python profiles/r2_probe.py --workload $wl --reps 3 --no-probe > gpurun_out/r2/ncu_$wl.log 2>&1
  tail -2 gpurun_out/r2/ncu_$wl.log
done
ls -la gpurun_out/r2
