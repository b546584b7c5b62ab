set -x
mkdir -p gpurun_out
python tools/instr_counts.py > gpurun_out/instr_counts.log 2>&1; tail -16 gpurun_out/instr_counts.log
cp gpurun_out/instr_counts_r2.json profiles/instr_counts_r2.json
python tools/config_bench.py > gpurun_out/configs_r2d.txt 2>&1
SCGPU_BENCH_IN_RANGE=1 python tools/config_bench.py > gpurun_out/configs_r2d_inrange.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2d.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2d_ref.json 2> gpurun_out/bench_r2d_ref.err
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2d_1gpu.json 2> gpurun_out/bench_r2d_1gpu.err; tail -c 400 gpurun_out/bench_r2d_1gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2d_1gpu.json'))
print(d['value'], d['roofline']['frac'], d['int_roofline']['warp_instr_per_product'], d['int_roofline']['frac'], d['checked_path']['value'], d['e2e']['value'])
for k,v in d['other_shapes'].items(): print(k, '%.4g'%v['per_s'], round(v['hbm_frac'],3))
r=json.load(open('gpurun_out/bench_r2d_ref.json')); print('ref', r['value'])
PY
