set -x
mkdir -p gpurun_out/r2
python profiles/r2_latency.py > gpurun_out/r2/latency.json 2> gpurun_out/r2/latency.err; tail -n 1 gpurun_out/r2/latency.json; tail -n 3 gpurun_out/r2/latency.err
