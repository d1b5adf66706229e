set -x
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29513"
python bench.py --no-cpu-baseline > gpurun_out/r2/bench_dino48_n1_s.json 2> gpurun_out/r2/bench_dino48_n1_s.err; tail -c 300 gpurun_out/r2/bench_dino48_n1_s.json
for n in 2 4 8; do $TR --nproc-per-node $n bench.py --gpus $n > gpurun_out/r2/bench_dino48_n${n}_s.json 2> gpurun_out/r2/bench_dino48_n${n}_s.err; tail -c 300 gpurun_out/r2/bench_dino48_n${n}_s.json; tail -2 gpurun_out/r2/bench_dino48_n${n}_s.err; done
$TR --nproc-per-node 8 bench.py --gpus 8 --workload dino_rounds --steps 5 --warmup 2 > gpurun_out/r2/bench_dino_rounds_n8.json 2> gpurun_out/r2/bench_dino_rounds_n8.err; tail -c 300 gpurun_out/r2/bench_dino_rounds_n8.json; tail -2 gpurun_out/r2/bench_dino_rounds_n8.err
