set -x
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4) 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/instr_counts.py > gpurun_out/instr_counts.log 2>&1; tail -16 gpurun_out/instr_counts.log | cut -c1-110
python tools/config_bench.py > gpurun_out/configs_r2e.txt 2>&1
SCGPU_BENCH_IN_RANGE=1 python tools/config_bench.py > gpurun_out/configs_r2e_inrange.txt 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2e_ref.json 2> gpurun_out/bench_r2e_ref.err
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2e_1gpu.json 2> gpurun_out/bench_r2e_1gpu.err; tail -c 400 gpurun_out/bench_r2e_1gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2e_1gpu.json'))
print(d['value'], d['roofline']['frac'], d['int_roofline']['warp_instr_per_product'], d['int_roofline']['frac'], d['checked_path']['value'], d['e2e']['value'])
for k,v in d['other_shapes'].items(): print(k, '%.4g'%v['per_s'], round(v['hbm_frac'],3))
r=json.load(open('gpurun_out/bench_r2e_ref.json')); print('ref', r['value'])
PY
