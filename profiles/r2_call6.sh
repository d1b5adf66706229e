set -x
mkdir -p gpurun_out/r2
nvidia-smi -L
python -m pytest tests/test_rounds_multi_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_multi_gpu_n2.log 2>&1; tail -8 gpurun_out/r2/pytest_multi_gpu_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 > gpurun_out/r2/bench_dino48_n2.json 2> gpurun_out/r2/bench_dino48_n2.err; tail -c 1500 gpurun_out/r2/bench_dino48_n2.json; tail -5 gpurun_out/r2/bench_dino48_n2.err
$TR bench.py --gpus 2 --workload ring128_1080p > gpurun_out/r2/bench_ring128_n2.json 2> gpurun_out/r2/bench_ring128_n2.err; tail -c 1200 gpurun_out/r2/bench_ring128_n2.json; tail -5 gpurun_out/r2/bench_ring128_n2.err
python bench.py --workload ring128_1080p --no-cpu-baseline > gpurun_out/r2/bench_ring128_n1.json 2> gpurun_out/r2/bench_ring128_n1.err; tail -c 800 gpurun_out/r2/bench_ring128_n1.json; tail -3 gpurun_out/r2/bench_ring128_n1.err
$TR bench.py --gpus 2 --workload dino_rounds --steps 5 --warmup 2 > gpurun_out/r2/bench_dino_rounds_n2.json 2> gpurun_out/r2/bench_dino_rounds_n2.err; tail -c 600 gpurun_out/r2/bench_dino_rounds_n2.json; tail -5 gpurun_out/r2/bench_dino_rounds_n2.err
