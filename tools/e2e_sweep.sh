# e2e pipeline shape sweep (diagnostic, run under gpurun): slots x reads per submitted batch
for cfg in ${SWEEP:-"3 131072" "4 131072" "6 65536" "4 262144"}; do set -- $cfg
python bench.py --no-cpu-baseline --zero-copy on --steps 5 --slots $1 --chunk $2 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('slots $1 chunk $2', {k: round(v['ms_per_step'],2) for k,v in d['e2e']['alternatives'].items()})"
done
