set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu -k "in_range or base_multiplication or fused or matvec" 2>&1 | tail -5
timeout 300 python tools/ab_basemul.py 2>&1 | tee gpurun_out/ab_basemul_r2.txt
