#!/bin/bash
# Quick GPU iteration (under gpurun): parity tests of the lift path, then the device-resident numbers of chr20 and wg.
# usage: bash tools/gpu_quick.sh <tag> [pytest -k expression]
TAG=${1:-x}
KEXPR=${2:-"parity or golden or fuzz"}
mkdir -p gpurun_out
echo "--- pytest"; timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -6
for WL in chr20 wg; do
  echo "--- bench $WL"
  timeout 900 python bench.py --workload $WL --no-assemble --no-cpu-baseline --zero-copy on 2>gpurun_out/bench_${TAG}_${WL}.err | tail -1 > gpurun_out/bench_${TAG}_${WL}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_${WL}.json"))
    print('value %.3e  ms/step %.3f  e2e %.3e (%.2f ms)  frac %.4f  stage_ms %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms']))
    print(d['config']['parity'][:120])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_${TAG}_${WL}.err").read()[-3000:])
PY
done
