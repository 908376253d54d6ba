timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/g16_pytest.log 2>&1; echo rc=$? >> gpurun_out/g16_pytest.log
tail -4 gpurun_out/g16_pytest.log
timeout 300 python tools/kernel_microbench.py --case A 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['case'], d['kernel'], d['ms'], d['frac_of_measured_hbm_peak'])
"
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/g16_bench.json 2> gpurun_out/g16_bench.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/g16_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:(round(v['ms'],3), round(v['frac'],3)) for k,v in d['kernels'].items()})
"
