mkdir -p gpurun_out
( timeout 2700 compute-sanitizer --tool memcheck --kernel-regex kns=scgpu python -m pytest tests -m gpu -q -x -k "not full_size and not at_scale and not bulk and not statistics and not os_entropy and not work_counter_batches" 2>&1 | tail -15 ) | tee gpurun_out/sanitizer_suite_r2c.txt
