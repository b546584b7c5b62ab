#!/bin/sh
# ncu --set full captures of the round-2 kernels (run on the GPU box: sh tools/capture_r2.sh); summaries land in gpurun_out/
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:k_polymul_w32 -s 3 -c 1 -o gpurun_out/polymul_r2 python tools/profile_run.py polymul 20 > /dev/null 2>&1
# final build of round 2: the headline configuration (plan with SCGPU_PLAN_INPUTS_IN_RANGE) and the shared-key product
$NCU -k regex:k_polymul_w32 -s 3 -c 1 -o gpurun_out/polymul_r2c python tools/profile_run.py polymul_inrange 20 > /dev/null 2>&1
$NCU -k regex:k_polymul_w32 -s 3 -c 1 -o gpurun_out/keyproduct_r2c python tools/profile_run.py keyproduct 20 > /dev/null 2>&1
$NCU -k regex:k_exact_w32 -s 1 -c 1 -o gpurun_out/exact_fwd_ref_r2 python tools/profile_run.py fwd 20 > /dev/null 2>&1
$NCU -k regex:k_exact_w32 -s 3 -c 1 -o gpurun_out/exact_inv_ref_r2 python tools/profile_run.py fwd 20 > /dev/null 2>&1
$NCU -k regex:k_exact_w32 -s 1 -c 1 -o gpurun_out/exact_fwd_avx_r2 python tools/profile_run.py exact_avx 20 > /dev/null 2>&1
$NCU -k regex:k_ber_lanes -s 1 -c 1 -o gpurun_out/ber_lanes_r2 python tools/profile_run.py ber 16 > /dev/null 2>&1
$NCU -k regex:k_stream_seq -s 1 -c 1 -o gpurun_out/ky64_r2 python tools/profile_run.py ky 16 > /dev/null 2>&1
$NCU -k regex:k_gen_rings -s 1 -c 1 -o gpurun_out/gen_rings_chacha_r2 python tools/profile_run.py randprod 16 > /dev/null 2>&1
$NCU -k regex:k_gen_rings -s 3 -c 1 -o gpurun_out/gen_rings_aes_r2 python tools/profile_run.py randprod 16 > /dev/null 2>&1
for f in polymul_r2 polymul_r2c keyproduct_r2c exact_fwd_ref_r2 exact_inv_ref_r2 exact_fwd_avx_r2 ber_lanes_r2 ky64_r2 gen_rings_chacha_r2 gen_rings_aes_r2; do
  python tools/ncu_summary.py gpurun_out/$f.ncu-rep gpurun_out/${f}_ncu.json > gpurun_out/${f}_summary.txt 2>&1
  rm -f gpurun_out/$f.ncu-rep
done
ls -la gpurun_out | tail -30
