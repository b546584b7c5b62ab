set -x
timeout 900 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu -k "base_multiplication or fused_polymul or proof_boundary or unaligned or work_counter" 2>&1 | tail -15
timeout 300 python tools/ab_basemul.py 2>&1 | tee gpurun_out/ab_basemul_r2.txt
