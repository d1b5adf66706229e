set -x
mkdir -p gpurun_out/r2
for wl in ring128_1080p ring256_4k; do
  MVS_K1_PAIRING=0 python profiles/r2_probe.py --workload $wl --no-probe > gpurun_out/r2/probe_${wl}_nopair.json 2>&1; tail -n 1 gpurun_out/r2/probe_${wl}_nopair.json
  python profiles/r2_probe.py --workload $wl --no-probe > gpurun_out/r2/probe_${wl}_pair.json 2>&1; tail -n 1 gpurun_out/r2/probe_${wl}_pair.json
done
