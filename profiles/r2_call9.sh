set -x
mkdir -p gpurun_out/r2
for per in 4 8 32 64; do MVS_K6_PER=$per python profiles/r2_probe.py --workload dino48 > gpurun_out/r2/probe_dino48_per$per.json 2>&1; tail -1 gpurun_out/r2/probe_dino48_per$per.json; done
python profiles/r2_probe.py --workload dino48 > gpurun_out/r2/probe_dino48_final.json 2>&1; tail -1 gpurun_out/r2/probe_dino48_final.json
for wl in temple47_mu5 temple47_mu7 ring256_4k; do python bench.py --workload $wl --steps 20 > gpurun_out/r2/bench_${wl}_n1.json 2> gpurun_out/r2/bench_${wl}_n1.err; tail -c 300 gpurun_out/r2/bench_${wl}_n1.json; tail -2 gpurun_out/r2/bench_${wl}_n1.err; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/bench_dino48_reference_arm.json 2> gpurun_out/r2/bench_reference.err; tail -c 600 gpurun_out/r2/bench_dino48_reference_arm.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2/plain_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/r2/launches_dino48.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu_launches.log 2>&1
python profiles/r2_probe.py --workload dino48 --reps 3 --no-probe > gpurun_out/r2/plain_dino48.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ncc_score -s 2 -c 1 -f -o gpurun_out/r2/ncu_dino48 python profiles/r2_probe.py --workload dino48 --reps 3 --no-probe > gpurun_out/r2/ncu_dino48.log 2>&1
ncu -i gpurun_out/r2/ncu_dino48.ncu-rep --page raw --csv > gpurun_out/r2/ncu_dino48_raw.csv 2>/dev/null
ncu -i gpurun_out/r2/ncu_dino48.ncu-rep --page source --csv > gpurun_out/r2/ncu_dino48_source.csv 2>/dev/null
rm -f gpurun_out/r2/ncu_dino48.ncu-rep
python bench.py > gpurun_out/r2/bench_dino48_n1_final.json 2> gpurun_out/r2/bench_dino48_n1_final.err; tail -c 300 gpurun_out/r2/bench_dino48_n1_final.json
