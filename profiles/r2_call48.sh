set -x
mkdir -p gpurun_out/r2
timeout 125 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/r2/launches_dino48_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu_launches_final.log 2>&1
tail -c 300 gpurun_out/r2/ncu_launches_final.log; wc -l gpurun_out/r2/launches_dino48_final.csv
