timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r27_launches_D.csv python bench.py --config D --batch 256 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --eager --profile-step > gpurun_out/r27_ncu.log 2>&1
python tools/summarise_ncu.py launches gpurun_out/r27_launches_D.csv | head -30
