mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_rand_product.py -x -q -m gpu -k "matvec or rand or module" 2>&1 | tail -3
timeout 300 python tools/ab_chunk.py 2>&1 | tee gpurun_out/ab_chunk_r2d.txt | grep "q8380417\|dilithium\|kyber"
