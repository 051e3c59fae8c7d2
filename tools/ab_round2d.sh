#!/bin/bash
# re-check of the Le = 512 cases and the decoder-attention cross-checks, then a source-level ncu capture of the persistent decode kernel
set -u
mkdir -p gpurun_out
O=gpurun_out
CASES="bwd_bf16_small_le512 bwd_bf16_small_le512_packed bwd_bf16_large_le512_packed bwd_bf16_small_le512_b6_packed \
xcheck_fbwd_dropout_small_le512 xcheck_fbwd_dropout_small_le512_packed xcheck_dattn_dropout_small xcheck_dattn_dropout_base_le256_packed \
bwd_bf16_small_ld12 gen_bf16_tiny gen_bf16_c5full bwd_bf16_base_le256_packed"
timeout 900 python tests/gpu_cases_model.py $CASES > $O/ab3_cases.log 2>&1
tail -1 $O/ab3_cases.log
grep -v '"ok": true' $O/ab3_cases.log | cut -c1-600 | head -8
NCU="ncu --profile-from-start off --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:decode_persistent_kernel -c 1 -f -o $O/p2b_decode python tools/profile_step.py eval > $O/p2b_ncu_decode.log 2>&1
ncu -i $O/p2b_decode.ncu-rep --page source --csv > $O/p2b_decode_source.csv 2>/dev/null
ncu -i $O/p2b_decode.ncu-rep --page details --csv > $O/p2b_decode_details.csv 2>/dev/null
rm -f $O/p2b_decode.ncu-rep
ls -la $O/p2b_decode*
