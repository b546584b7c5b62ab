set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:k_polymul_w32 -s 3 -c 1 -o gpurun_out/polymul_r2bm python tools/profile_run.py polymul 20 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/polymul_r2bm.ncu-rep gpurun_out/polymul_r2bm_ncu.json > gpurun_out/polymul_r2bm_summary.txt 2>&1
rm -f gpurun_out/polymul_r2bm.ncu-rep
