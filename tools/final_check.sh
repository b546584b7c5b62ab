mkdir -p gpurun_out
timeout 300 python tools/dropin_latency.py 2>&1 | tee gpurun_out/dropin_latency_r2.txt | tail -16
