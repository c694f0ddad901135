#!/bin/bash
# One GPU iteration: parity tests, the bench line, and an ncu capture of the dominant kernel.
# usage (under gpurun): bash tools/gpu_cycle.sh <tag> [kernel-regex]
TAG=${1:-x}
KRE=${2:-lift_pairs}
mkdir -p gpurun_out
echo "--- pytest"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "--- bench"; timeout 900 python bench.py 2>gpurun_out/bench_${TAG}.err | tail -1 > gpurun_out/bench_${TAG}.json; cat gpurun_out/bench_${TAG}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.3e  ms/step %.3f  e2e %.3e (%.2f ms)  frac %.4f  stage_ms %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms']))
print('e2e alt', d['e2e']['alternatives']); print('cpu', d['cpu_baseline']); print('clocks', d['clocks'], 'launches', d['gpu_launches'])
" || tail -5 gpurun_out/bench_${TAG}.err
echo "--- ncu full ($KRE)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 1 -c 1 -f -o gpurun_out/prof_${TAG} python bench.py --reads 200000 --steps 1 --warmup 1 --no-cpu-baseline --zero-copy off > gpurun_out/ncu_${TAG}.log 2>&1; tail -2 gpurun_out/ncu_${TAG}.log | cut -c1-300
