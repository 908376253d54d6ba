for d in 0 8; do echo "debug=$d"; SPAIR_GEMM_DEBUG=$d timeout 200 python tools/gemm_check.py --decoder-only 2>&1 | grep -v Warn; done
