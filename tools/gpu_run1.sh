set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_dropin.py -x -q -m gpu -k "exact or golden or ibe or members or config_c1 or divides" 2>&1 | tail -4
python tools/config_bench.py 2>&1 | grep "exact.*inv\|C4 exact inv" | tee gpurun_out/configs_exact_inv_r2d.txt
