set -x
timeout 900 python -m pytest tests/test_gpu_gauss.py -x -q -m gpu -k "high_precision or dropin_sampler" 2>&1 | tail -15
