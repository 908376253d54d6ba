timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/g13_pytest.log 2>&1; echo rc=$? >> gpurun_out/g13_pytest.log
tail -6 gpurun_out/g13_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/g13_bench.json 2> gpurun_out/g13_bench.err; tail -3 gpurun_out/g13_bench.err | cut -c1-200; python -c "
import json
d=json.loads([l for l in open('gpurun_out/g13_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
"
