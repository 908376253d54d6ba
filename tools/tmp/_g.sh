python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r22_pytest.log 2>&1; echo rc=$? >> gpurun_out/r22_pytest.log; tail -3 gpurun_out/r22_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r22_bench.json 2> gpurun_out/r22_bench.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r22_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['traffic'])
"
