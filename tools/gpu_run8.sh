mkdir -p gpurun_out
timeout 900 python tools/fuzz_parity.py 300 20261017 2>&1 | tail -12 | tee gpurun_out/fuzz_r2c.txt
