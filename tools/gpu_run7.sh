set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r2c_2gpu.json 2> gpurun_out/bench_r2c_2gpu.err; tail -c 300 gpurun_out/bench_r2c_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_r2c_ref2.json 2> gpurun_out/bench_r2c_ref2.err; cat gpurun_out/bench_r2c_ref2.json | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2c_2gpu.json'))
print(d['value'], d['n_gpus'], d['e2e']['value'], d.get('e2e_multi'))
PY
