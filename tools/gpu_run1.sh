set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rand_product.py tests/test_gpu_ntt.py -x -q -m gpu -k "rand or fused_polymul" 2>&1 | tail -5
timeout 600 python tools/ab_randprod.py 2>&1 | tee gpurun_out/ab_randprod_r2.txt
timeout 200 python tools/fuzz_parity.py 60 7 2>&1 | tail -2
