timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/g11_pytest.log 2>&1; echo rc=$? >> gpurun_out/g11_pytest.log
tail -4 gpurun_out/g11_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/g11_bench.json 2> gpurun_out/g11_bench.err; tail -3 gpurun_out/g11_bench.err | cut -c1-200; python -c "
import json
d=json.loads([l for l in open('gpurun_out/g11_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); print({k:(round(v['ms'],3), round(v['frac'],3)) for k,v in d['kernels'].items()})
"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g11_launches.csv python bench.py --steps 2 --warmup 4 --no-cpu-baseline --no-extras --eager --profile-step > gpurun_out/g11_ncu_bench.log 2>&1
python tools/summarise_ncu.py launches gpurun_out/g11_launches.csv | head -40
