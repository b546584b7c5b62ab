set -x
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5) 2>&1
python tools/instr_counts.py > gpurun_out/instr_counts.log 2>&1; tail -3 gpurun_out/instr_counts.log
cp profiles/instr_counts_r2.json gpurun_out/instr_counts_r2.json
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2b_1gpu.json 2> gpurun_out/bench_r2b_1gpu.err; tail -c 600 gpurun_out/bench_r2b_1gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2b_1gpu.json'))
print(d['value'], d['roofline']['frac'], d['int_roofline'], d['checked_path'], d['e2e']['value'])
PY
