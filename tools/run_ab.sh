mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3) 2>&1 | tee gpurun_out/gputests_zc.log
H=tests/harness/build/table_harness
timeout 120 $H libsafecrypto_b200/libscgpu.so oracle/_ref/libscref.so --rounds 1 --time 2000 2>&1 | grep "TIME\|SUMMARY" | tee gpurun_out/dropin_table_latency_r2.txt
SCGPU_DROPIN_ZERO_COPY=0 timeout 120 $H libsafecrypto_b200/libscgpu.so oracle/_ref/libscref.so --rounds 1 --time 2000 2>&1 | grep "TIME\|SUMMARY" | sed 's/^/copies: /' | tee -a gpurun_out/dropin_table_latency_r2.txt
timeout 250 python tools/ab_canonical2.py 2>&1 > gpurun_out/ab_canonical2_chunk2.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2g_1gpu.json 2> gpurun_out/bench_r2g_1gpu.err; tail -c 300 gpurun_out/bench_r2g_1gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2g_1gpu.json'))
print(d['value'], d['roofline']['frac'], d['checked_path']['value'], d['e2e']['value'])
for k,v in d['other_shapes'].items(): print(k, round(v['hbm_frac'],3))
PY
