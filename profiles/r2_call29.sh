set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_pmvs_gpu.py tests/test_torch_ops.py -m gpu -x -q > gpurun_out/r2/pytest_gpu_call29.log 2>&1; tail -n 3 gpurun_out/r2/pytest_gpu_call29.log
for wl in temple47_mu5 temple47_mu7; do python bench.py --workload $wl --steps 20 > gpurun_out/r2/bench_${wl}_n1.json 2> gpurun_out/r2/bench_${wl}_n1.err; tail -c 300 gpurun_out/r2/bench_${wl}_n1.json; tail -n 2 gpurun_out/r2/bench_${wl}_n1.err; done
for wl in temple47_mu5 temple47_mu7; do
  python profiles/r2_probe.py --workload $wl --reps 3 --no-probe > gpurun_out/r2/plain_$wl.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:ncc_score -s 2 -c 1 -f -o gpurun_out/r2/ncu_$wl python profiles/r2_probe.py --workload $wl --reps 3 --no-probe > gpurun_out/r2/ncu_$wl.log 2>&1
  ncu -i gpurun_out/r2/ncu_$wl.ncu-rep --page raw --csv > gpurun_out/r2/ncu_${wl}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2/ncu_$wl.ncu-rep --page source --csv > gpurun_out/r2/ncu_${wl}_source.csv 2>/dev/null
  rm -f gpurun_out/r2/ncu_$wl.ncu-rep
done
