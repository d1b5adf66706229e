set -x
mkdir -p gpurun_out/r2
python -m pytest tests/test_rounds_gpu.py tests/test_rounds_multi_gpu.py tests/test_pmvs_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_multi_gpu_n2.log 2>&1; tail -n 4 gpurun_out/r2/pytest_multi_gpu_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 100 > gpurun_out/r2/bench_dino48_n2_c.json 2> gpurun_out/r2/bench_dino48_n2_c.err; tail -c 400 gpurun_out/r2/bench_dino48_n2_c.json; tail -n 3 gpurun_out/r2/bench_dino48_n2_c.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --workload dino_rounds --steps 10 --warmup 2 > gpurun_out/r2/bench_dino_rounds_n2.json 2> gpurun_out/r2/bench_dino_rounds_n2.err; tail -c 400 gpurun_out/r2/bench_dino_rounds_n2.json; tail -n 3 gpurun_out/r2/bench_dino_rounds_n2.err
