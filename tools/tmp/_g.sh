timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r26_pytest.log 2>&1; echo rc=$? >> gpurun_out/r26_pytest.log; tail -3 gpurun_out/r26_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('A', d['value'], d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['kernels'].items() if 'sweep' in k})
"
for c in "C 64" "C 512" "D 256"; do set -- $c; timeout 300 python bench.py --config $1 --batch $2 --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$c', d['value'], d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['kernels'].items() if 'sweep' in k})
"; done
