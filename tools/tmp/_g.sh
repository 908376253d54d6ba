timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r24_pytest.log 2>&1; echo rc=$? >> gpurun_out/r24_pytest.log; tail -3 gpurun_out/r24_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r24_bench.json 2> gpurun_out/r24_bench.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r24_bench.json') if l.startswith('{')][-1])
print('A', d['value'], d['ms_per_step'], d['e2e']['value'])
"
for b in 64 512; do timeout 300 python bench.py --config C --batch $b --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r24_C$b.json 2>/dev/null; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r24_C$b.json') if l.startswith('{')][-1])
print('C $b', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
"; done
