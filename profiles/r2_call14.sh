set -x
mkdir -p gpurun_out/r2
P=simple-implementation-of-structure-from-motion-and-multi-view-stereo-by-python_b200
python -m pytest tests/test_pmvs_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_pmvs_a.log 2>&1; tail -3 gpurun_out/r2/pytest_pmvs_a.log
MVS_LIB=$PWD/$P/libmvsncc_magicfloor.so python -m pytest tests/test_pmvs_gpu.py -m gpu -x -q > gpurun_out/r2/pytest_pmvs_magic.log 2>&1; tail -3 gpurun_out/r2/pytest_pmvs_magic.log
for wl in temple47_mu5 temple47_mu7; do
  python profiles/r2_probe.py --workload $wl --no-probe > gpurun_out/r2/k2_${wl}_10f.json 2>&1; tail -1 gpurun_out/r2/k2_${wl}_10f.json
  MVS_LIB=$PWD/$P/libmvsncc_magicfloor.so python profiles/r2_probe.py --workload $wl --no-probe > gpurun_out/r2/k2_${wl}_10f_magic.json 2>&1; tail -1 gpurun_out/r2/k2_${wl}_10f_magic.json
done
