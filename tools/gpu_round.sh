#!/bin/bash
# One full GPU evidence pass (under gpurun): parity tests, smoke, the bench line of both arms, the ncu launch list of the bench
# command, and an ncu --set full capture of the dominant kernel on the bench's own launch shape.  usage: bash tools/gpu_round.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
echo "--- pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "--- smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "--- bench"; timeout 1500 python bench.py 2>gpurun_out/bench_${TAG}.err | tail -1 > gpurun_out/bench_${TAG}.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print('value %.3e  ms/step %.3f  e2e %.3e (%.2f ms)  with_pack %.3e  frac %.4f  stage_ms %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_with_pack']['value'], d['roofline']['frac'], d['roofline']['stage_ms']))
print(d['config']['parity']); print('cpu', d['cpu_baseline']); print('clocks', d['clocks'], 'launches', d['gpu_launches'])
for k in ('assemble_bases', 'assemble_records'):
    print(k, d[k]['kernel_ms'], d[k]['roofline']['frac'])
print('frame_records', d['assemble_records']['frame_records']['kernel_ms'], 'e2e_records', d['e2e_records']['records_per_s'], d['e2e_records']['pcie_gbs'])
PY
echo "--- reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_${TAG}_reference.json; cut -c1-300 gpurun_out/bench_${TAG}_reference.json
echo "--- ncu launch list"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-assemble --no-parity --zero-copy on > gpurun_out/launches_${TAG}.log 2>&1; tail -1 gpurun_out/launches_${TAG}.log | cut -c1-200
if [ -n "$SKIP_FULL" ]; then exit 0; fi
echo "--- ncu full: lift_pairs on the bench launch shape (wg, 2 M reads per launch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lift_pairs_kernel" -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_lift python tools/variant_times.py wg 2000000 > gpurun_out/ncu_${TAG}_lift.log 2>&1; tail -1 gpurun_out/ncu_${TAG}_lift.log | cut -c1-200
