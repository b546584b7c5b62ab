set -x
mkdir -p gpurun_out
python tools/config_bench.py > gpurun_out/configs_r2c.txt 2>&1; tail -3 gpurun_out/configs_r2c.txt
SCGPU_BENCH_IN_RANGE=1 python tools/config_bench.py > gpurun_out/configs_r2c_inrange.txt 2>&1; tail -3 gpurun_out/configs_r2c_inrange.txt
{
echo "r2c: compute-sanitizer on tools/sanitize_run.py (round-2 kernels: base-multiplication polymul, k_key_residues + residue-table key product, per-row keys, kernels without range votes, pipelined work-counter claim, canonical transforms, mat-vec, samplers)"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run done" | sed "s/^/  $tool: /"
done
} > gpurun_out/sanitizer_r2c.txt 2>&1
cat gpurun_out/sanitizer_r2c.txt
