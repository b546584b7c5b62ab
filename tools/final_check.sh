set -x
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3) 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/instr_counts.py > gpurun_out/instr_counts.log 2>&1
python tools/config_bench.py > gpurun_out/configs_r2g.txt 2>&1
SCGPU_BENCH_IN_RANGE=1 python tools/config_bench.py > gpurun_out/configs_r2g_inrange.txt 2>&1
python tools/gauss_bench.py 18 512 18 > gpurun_out/gauss_r2g.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2g.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2g_ref.json 2> gpurun_out/bench_r2g_ref.err
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2g_1gpu.json 2> gpurun_out/bench_r2g_1gpu.err; tail -c 300 gpurun_out/bench_r2g_1gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2g_1gpu.json'))
print(d['value'], d['roofline']['frac'], d['int_roofline']['warp_instr_per_product'], d['checked_path']['value'], d['e2e']['value'])
g=d['gaussian']
for k,v in g.items():
    if isinstance(v,dict) and 'samples_per_s' in v: print(k, '%.3g'%v['samples_per_s'])
r=json.load(open('gpurun_out/bench_r2g_ref.json')); print('ref', r['value'])
PY
tests/harness/build/table_harness libsafecrypto_b200/libscgpu.so oracle/_ref/libscref.so --rounds 1 --time 2000 2>&1 | grep "TIME\|SUMMARY" > gpurun_out/dropin_table_latency_r2g.txt
