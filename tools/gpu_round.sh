#!/bin/bash
# One full GPU evidence pass (under gpurun): parity tests, the bench line, the ncu launch list of the same command, and
# ncu --set full captures of the dominant kernels.  usage: bash tools/gpu_round.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
echo "--- pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "--- smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "--- bench"; timeout 1200 python bench.py 2>gpurun_out/bench_${TAG}.err | tail -1 > gpurun_out/bench_${TAG}.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print('value %.3e  ms/step %.3f  e2e %.3e (%.2f ms)  frac %.4f  stage_ms %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms']))
print('cpu', d['cpu_baseline']); print('clocks', d['clocks'], 'launches', d['gpu_launches'])
for k in ('assemble_bases', 'assemble_records'):
    print(k, d[k]['kernel_ms'], d[k]['roofline'], d[k]['cpu_baseline']['value'])
PY
echo "--- reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
echo "--- ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${TAG}.log 2>&1; tail -1 gpurun_out/launches_${TAG}.log | cut -c1-200
echo "--- ncu full: lift_pairs, warp_pairs, emit_records (1M reads)"; if [ -n "$SKIP_FULL" ]; then exit 0; fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lift_pairs|warp_pairs|emit_records" -s 3 -c 3 -f -o gpurun_out/prof_${TAG}_lift python tools/variant_times.py chr20 1000000 > gpurun_out/ncu_${TAG}_lift.log 2>&1; tail -1 gpurun_out/ncu_${TAG}_lift.log | cut -c1-200
echo "--- ncu full: bam_write"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bam_write" -s 4 -c 2 -f -o gpurun_out/prof_${TAG}_bam python bench.py --no-cpu-baseline --steps 2 > gpurun_out/ncu_${TAG}_bam.log 2>&1; tail -1 gpurun_out/ncu_${TAG}_bam.log | cut -c1-200
