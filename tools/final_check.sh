set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu -k "in_range or canonical or fused_polymul" 2>&1 | tail -4
python tools/config_bench.py 2>&1 | grep "canonical"
SCGPU_BENCH_IN_RANGE=1 python tools/config_bench.py 2>&1 | grep "canonical" | sed 's/^/in-range: /'
