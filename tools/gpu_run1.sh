set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu -k "key or in_range or work_counter or full_size" 2>&1 | tail -5
timeout 300 python tools/ab_basemul.py 2>&1 | grep "key product"
