set -x
mkdir -p gpurun_out/r2
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29512"
$TR --nproc-per-node 8 bench.py --gpus 8 > gpurun_out/r2/bench_dino48_n8.json 2> gpurun_out/r2/bench_dino48_n8.err; tail -c 400 gpurun_out/r2/bench_dino48_n8.json; tail -3 gpurun_out/r2/bench_dino48_n8.err
$TR --nproc-per-node 4 bench.py --gpus 4 > gpurun_out/r2/bench_dino48_n4.json 2> gpurun_out/r2/bench_dino48_n4.err; tail -c 400 gpurun_out/r2/bench_dino48_n4.json; tail -3 gpurun_out/r2/bench_dino48_n4.err
python bench.py --no-cpu-baseline > gpurun_out/r2/bench_dino48_n1_box8.json 2> gpurun_out/r2/bench_dino48_n1_box8.err; tail -c 400 gpurun_out/r2/bench_dino48_n1_box8.json
$TR --nproc-per-node 8 bench.py --gpus 8 --workload ring128_1080p > gpurun_out/r2/bench_ring128_n8.json 2> gpurun_out/r2/bench_ring128_n8.err; tail -c 400 gpurun_out/r2/bench_ring128_n8.json; tail -3 gpurun_out/r2/bench_ring128_n8.err
python bench.py --workload ring128_1080p --no-cpu-baseline > gpurun_out/r2/bench_ring128_n1_box8.json 2> gpurun_out/r2/bench_ring128_n1_box8.err; tail -c 400 gpurun_out/r2/bench_ring128_n1_box8.json
$TR --nproc-per-node 4 bench.py --gpus 4 --workload ring128_1080p > gpurun_out/r2/bench_ring128_n4.json 2> gpurun_out/r2/bench_ring128_n4.err; tail -c 400 gpurun_out/r2/bench_ring128_n4.json
$TR --nproc-per-node 8 bench.py --gpus 8 --workload ring256_4k --steps 20 > gpurun_out/r2/bench_ring256_n8.json 2> gpurun_out/r2/bench_ring256_n8.err; tail -c 400 gpurun_out/r2/bench_ring256_n8.json; tail -3 gpurun_out/r2/bench_ring256_n8.err
python -m pytest tests/test_rounds_multi_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_multi_gpu_n8box.log 2>&1; tail -5 gpurun_out/r2/pytest_multi_gpu_n8box.log
