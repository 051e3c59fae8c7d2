#!/bin/bash
# Round-2 evidence run on ONE B200 (under gpurun).  Everything lands in gpurun_out/; tools/summarise_profiles.py r02
# condenses it into profiles/.  Numbers printed under ncu are never used as bench values.
#   1. launch lists (gpu__time_duration) of one train step and one eval batch
#   2. one `ncu --set full` capture of every kernel the north star names that round 1 had not captured
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
NCU="ncu --profile-from-start off --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_train_step_launches.csv python tools/profile_step.py > gpurun_out/${TAG}_ncu_train.log 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_eval_batch_launches.csv python tools/profile_step.py eval > gpurun_out/${TAG}_ncu_eval.log 2>&1
# train-step kernels (first match after 3 warm-up steps; -c 1: one launch, ~40 replays)
for K in "gemm_tc_kernel:gemm" "fattn_fwd_kernel:fattn_fwd" "fattn_bwd_kernel:fattn_bwd" "adamw_kernel:adamw" "ce_fwd_kernel:ce_fwd" "ce_bwd_kernel:ce_bwd" "dattn_fwd_kernel:dattn_fwd" "dattn_bwd_kernel:dattn_bwd" "embed_fwd_kernel:embed_fwd" "embed_bwd_kernel:embed_bwd" "rmsnorm_fwd:rmsnorm_fwd" "rmsnorm_bwd:rmsnorm_bwd" "sumsq_partial_kernel:sumsq"; do
  PAT=${K%%:*}; NAME=${K##*:}
  timeout 240 $NCU --set full --import-source on -k regex:$PAT -c 1 -o gpurun_out/${TAG}_${NAME} python tools/profile_step.py > gpurun_out/${TAG}_ncu_${NAME}.log 2>&1
done
# eval: the persistent decode kernel (one launch = the whole beam search of a batch)
timeout 600 $NCU --set full --import-source on -k regex:decode_persistent_kernel -c 1 -o gpurun_out/${TAG}_decode_persistent python tools/profile_step.py eval > gpurun_out/${TAG}_ncu_decode.log 2>&1
# summarise on the box (gpurun_out/ comes back only while it stays under 64 MiB) and keep the two reports worth reading at source level
P5_PROF_DIR=gpurun_out/profiles_${TAG} python tools/summarise_profiles.py ${TAG}
ls -la gpurun_out/*.ncu-rep | tail -20
for f in gpurun_out/${TAG}_*.ncu-rep; do case "$f" in *_gemm.ncu-rep|*_decode_persistent.ncu-rep) ;; *) rm -f "$f";; esac; done
du -sh gpurun_out
