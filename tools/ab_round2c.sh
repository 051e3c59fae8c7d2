#!/bin/bash
# attention-kernel changes (hash key hoisted, compile-time dropout / full-slice variants, table work on two warps):
# parity on every case that exercises them, then the train bench and an ncu launch list for the per-kernel times
set -u
mkdir -p gpurun_out
O=gpurun_out
CASES="bwd_bf16_base_le256 bwd_bf16_base_le256_packed bwd_bf16_small_le512 bwd_bf16_small_le512_packed bwd_bf16_large_le512_packed \
bwd_bf16_small_le512_b6_packed xcheck_fbwd_dropout_small xcheck_fbwd_dropout_base_le256_packed xcheck_fbwd_dropout_base_le256 \
xcheck_fbwd_dropout_small_le512 xcheck_fbwd_dropout_small_le512_packed dropout_bf16_small bwd_bf16_small bwd_bf16_small_packed \
bwd_bf16_tiny bwd_bf16_tiny_b1 bwd_bf16_tiny_le8 bwd_bf16_c2full_packed bwd_bf16_c2full xcheck_dattn_dropout_small \
xcheck_dattn_dropout_base_le256_packed bwd_bf16_small_ld12 gen_bf16_tiny gen_bf16_c5full gen_bf16_v32600 bwd_bf16_gated_small"
timeout 900 python tests/gpu_cases_model.py $CASES > $O/ab2_cases.log 2>&1
tail -1 $O/ab2_cases.log
grep -v '"ok": true' $O/ab2_cases.log | head -20
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference"
timeout 300 $B > $O/ab2_bench.json 2> $O/ab2_bench.err
timeout 300 $B > $O/ab2_bench_b.json 2> $O/ab2_bench_b.err
timeout 300 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum --csv --log-file $O/ab2_train_launches.csv python tools/profile_step.py > $O/ab2_ncu.log 2>&1
python tools/agg_launches.py $O/ab2_train_launches.csv 2>/dev/null | head -14
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab2_bench*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        ev = d.get("eval") or {}
        print("%-30s train %8.1f samples/s %7.3f ms/step  gemm256 %.0f TF/s  eval %.0f items/s %.3f ms/batch" % (
            f.split("/")[-1], d["value"], d["ms_per_step"], d["roofline"]["achieved"], ev["value"], ev["ms_per_batch"]))
    except Exception as e:
        print(f, "failed", repr(e)[:200])
PY
