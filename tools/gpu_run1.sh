set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_rand_product.py -x -q -m gpu -k "matvec or rand or in_range" 2>&1 | tail -4
timeout 300 python tools/ab_basemul.py 2>&1 | grep "mat-vec"
timeout 200 python tools/fuzz_parity.py 45 13 2>&1 | tail -2
