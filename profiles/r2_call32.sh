set -x
mkdir -p gpurun_out/r2
python profiles/pcie_probe.py > gpurun_out/r2/pcie_probe.log 2>&1; tail -n 6 gpurun_out/r2/pcie_probe.log
for ch in 32768 65536 131072 262144 524288; do
  MVS_HOST_CHUNK=$ch python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2/e2e_chunk_$ch.json 2> gpurun_out/r2/e2e_chunk_$ch.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2/e2e_chunk_$ch.json').read().strip().splitlines()[-1]); print($ch, d['ms_per_step'], d['e2e']['value'])"
done
