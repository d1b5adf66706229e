set -x
mkdir -p gpurun_out/r2
for u in 0 1 2 3; do
  MVS_K1_UNR=$u python profiles/r2_probe.py --workload ring128_1080p --reps 5 --no-probe > gpurun_out/r2/k1_unr${u}_ring128.log 2>&1; tail -n 1 gpurun_out/r2/k1_unr${u}_ring128.log
done
for u in 0 2; do
  MVS_K1_UNR=$u python profiles/r2_probe.py --workload ring256_4k --reps 3 --no-probe > gpurun_out/r2/k1_unr${u}_ring256.log 2>&1; tail -n 1 gpurun_out/r2/k1_unr${u}_ring256.log
done
