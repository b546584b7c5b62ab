set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu -k "work_counter or fused_polymul or full_size or canonical or divides" 2>&1 | tail -4
timeout 300 python tools/ab_basemul.py 2>&1 | tee gpurun_out/ab_basemul_r2.txt
timeout 600 python tools/ab_sched.py 2>&1 | tee gpurun_out/ab_sched_r2.txt | head -12
