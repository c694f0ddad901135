for R in 750000 1500000; do for V in 1 2 3; do
python bench.py --reads $R --value-slots $V --no-assemble --no-cpu-baseline --zero-copy on --no-parity 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('reads $R V $V value %.3e ms %.4f e2e %.3e ms %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']))"
done; done
