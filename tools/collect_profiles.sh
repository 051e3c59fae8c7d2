#!/bin/bash
# Evidence run on ONE B200 (under gpurun): default bench (with CPU baseline), ncu launch lists of one train step and one
# eval batch, and one `ncu --set full` capture of each hot kernel.  Everything lands in gpurun_out/; summaries are
# extracted afterwards with tools/summarise_profiles.py (numbers printed under ncu are never used as bench values).
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.log 2>&1
NCU="ncu --profile-from-start off --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_train_step_launches.csv python tools/profile_step.py > gpurun_out/${TAG}_ncu_train.log 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_eval_batch_launches.csv python tools/profile_step.py eval > gpurun_out/${TAG}_ncu_eval.log 2>&1
for K in "gemm_tc_kernel:gemm" "fattn_fwd_kernel:fattn_fwd" "fattn_bwd_kernel:fattn_bwd" "dattn_bwd_kernel:dattn_bwd" "rmsnorm_bwd_vec:rmsnorm_bwd" "adamw_kernel:adamw"; do
  PAT=${K%%:*}; NAME=${K##*:}
  timeout 300 $NCU --set full --import-source on -k regex:$PAT -s 3 -c 2 -o gpurun_out/${TAG}_${NAME} python tools/profile_step.py > gpurun_out/${TAG}_ncu_${NAME}.log 2>&1
done
tail -1 gpurun_out/${TAG}_bench_n1.log | cut -c1-400
