set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu -k "base_multiplication or fused or proof_boundary or unaligned or work_counter or canonical or matvec or full_size" 2>&1 | tail -5
timeout 300 python tools/ab_basemul.py 2>&1 | tee gpurun_out/ab_basemul_r2.txt
timeout 600 python tools/ab_sched.py 2>&1 | tee gpurun_out/ab_sched_r2.txt
