set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:k_polymul_w32 -s 3 -c 1 -o gpurun_out/polymul_r2c python tools/profile_run.py polymul_inrange 20 > /dev/null 2>&1
$NCU -k regex:k_polymul_w32 -s 3 -c 1 -o gpurun_out/keyproduct_r2c python tools/profile_run.py keyproduct 20 > /dev/null 2>&1
for f in polymul_r2c keyproduct_r2c; do
  python tools/ncu_summary.py gpurun_out/$f.ncu-rep gpurun_out/${f}_ncu.json > gpurun_out/${f}_summary.txt 2>&1
  rm -f gpurun_out/$f.ncu-rep
done
python tools/instr_counts.py > gpurun_out/instr_counts.log 2>&1; tail -16 gpurun_out/instr_counts.log
cp gpurun_out/instr_counts_r2.json profiles/instr_counts_r2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2c.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2c_1gpu.json 2> gpurun_out/bench_r2c_1gpu.err; tail -c 400 gpurun_out/bench_r2c_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2c_ref.json 2> gpurun_out/bench_r2c_ref.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2c_1gpu.json'))
print(d['value'], d['roofline']['frac'], d['int_roofline']['warp_instr_per_product'], d['int_roofline']['frac'], d['checked_path']['value'], d['e2e']['value'])
for k,v in d['other_shapes'].items(): print(k, '%.4g'%v['per_s'], round(v['hbm_frac'],3))
r=json.load(open('gpurun_out/bench_r2c_ref.json')); print('ref', r['value'])
PY
