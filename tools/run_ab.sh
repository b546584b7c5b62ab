mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
O=gpurun_out/sanitizer_r2g.txt
echo "r2g: compute-sanitizer on tools/sanitize_run.py (adds: chunked work-counter claims K = 2, 4 over a batch with both phases) and on the drop-in table harness (count = 1 calls on mapped host memory)" > $O
for tool in memcheck racecheck synccheck; do
  timeout 900 $CS --tool $tool python tools/sanitize_run.py 2>&1 | grep "parity\|SUMMARY" | sed "s/^/  $tool: /" >> $O
done
timeout 600 $CS --tool memcheck tests/harness/build/table_harness libsafecrypto_b200/libscgpu.so oracle/_ref/libscref.so --rounds 1 --threads 4 2>&1 | grep "SUMMARY" | sed "s/^/  table harness memcheck: /" >> $O
cat $O
timeout 400 python tools/fuzz_parity.py 300 20261018 2>&1 | tail -4 | tee gpurun_out/fuzz_r2g.txt
