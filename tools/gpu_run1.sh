set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/config_bench.py 2>&1 | grep "exact" > gpurun_out/configs_exact_r2e.txt; cat gpurun_out/configs_exact_r2e.txt | cut -c1-46,74-100
