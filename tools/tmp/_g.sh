SPAIR_NVCC_EXTRA=-DSW_TIMING python spair_pytorch_b200/_build.py --force > /dev/null 2>&1
echo "== C 64"; python tools/sweep_phase_timing.py C 64 2>&1 | grep -v Warn | tail -30
echo "== A 256"; python tools/sweep_phase_timing.py A 256 2>&1 | grep -v Warn | tail -30
