#!/bin/bash
# e2e pipeline shape sweep (under gpurun): slots x chunk on the default workload.  usage: bash tools/e2e_sweep.sh
for CFG in "3 131072" "4 131072" "6 131072" "4 65536" "6 65536" "4 262144"; do set -- $CFG
python bench.py --slots $1 --chunk $2 --no-assemble --no-cpu-baseline --zero-copy on --no-parity 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('slots $1 chunk $2 value %.3e e2e %.3e ms %.3f floor %.3f with_pack %.3e'%(d['value'],e['value'],e['ms_per_step'],e['copy_only']['ms_per_step'],d['e2e_with_pack']['value']))"
done
