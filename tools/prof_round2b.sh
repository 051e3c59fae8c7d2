#!/bin/bash
# source-level ncu captures of the attention kernels and the small decoder GEMM (one launch each), exported as CSV pages
set -u
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
for K in "fattn_fwd_kernel:fattn_fwd" "fattn_bwd_kernel:fattn_bwd" "rmsnorm_bwd:rmsnorm_bwd"; do
  PAT=${K%%:*}; NAME=${K##*:}
  timeout 240 $NCU --set full --import-source on -k regex:$PAT -c 1 -f -o gpurun_out/p2b_${NAME} python tools/profile_step.py > gpurun_out/p2b_ncu_${NAME}.log 2>&1
  ncu -i gpurun_out/p2b_${NAME}.ncu-rep --page source --csv > gpurun_out/p2b_${NAME}_source.csv 2>/dev/null
  ncu -i gpurun_out/p2b_${NAME}.ncu-rep --page details --csv > gpurun_out/p2b_${NAME}_details.csv 2>/dev/null
done
# the decoder's M = 512 GEMM chain: skip the encoder launches (launch-skip) and take one bn64 launch
timeout 240 $NCU --set full --import-source on -k regex:gemm_tc_kernel -s 60 -c 1 -f -o gpurun_out/p2b_gemm64 python tools/profile_step.py > gpurun_out/p2b_ncu_gemm64.log 2>&1
ncu -i gpurun_out/p2b_gemm64.ncu-rep --page source --csv > gpurun_out/p2b_gemm64_source.csv 2>/dev/null
ncu -i gpurun_out/p2b_gemm64.ncu-rep --page details --csv > gpurun_out/p2b_gemm64_details.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/p2b_*.ncu-rep
du -sh gpurun_out
