#!/bin/bash
# Round-2 (second session) A/B under gpurun, one GPU:
#   1. parity of the new code paths: fused attention backward at 256 < Le <= 512, early gradient norm (every case that goes
#      through train_step), plus two full-size regressions
#   2. bench A/B: early gradient norm off / on, tile-width cost model, configs[3] with / without the Le = 512 fused backward
set -u
mkdir -p gpurun_out
O=gpurun_out
CASES="bwd_bf16_small_le512 bwd_bf16_small_le512_packed bwd_bf16_large_le512_packed bwd_bf16_small_le512_b6_packed \
xcheck_fbwd_dropout_small_le512 xcheck_fbwd_dropout_small_le512_packed bwd_bf16_base_le256_packed xcheck_fbwd_dropout_base_le256_packed \
adamw_fp32_tiny adamw_fp32_tiny_packed adamw_fp32_tiny_async asyncopt_bf16_small_bitwise adamw_x3_tiny resume_fp32_tiny \
resume_bf16_small varlen_bf16_small bwd_bf16_c2full_packed"
timeout 900 python tests/gpu_cases_model.py $CASES > $O/ab_cases.log 2>&1
tail -1 $O/ab_cases.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference"
P5_NO_EARLY_NORM=1 timeout 300 $B > $O/ab_bench_noearly.json 2> $O/ab_bench_noearly.err
timeout 300 $B > $O/ab_bench_default.json 2> $O/ab_bench_default.err
P5_TILE_MODEL=1 timeout 300 $B > $O/ab_bench_tilemodel.json 2> $O/ab_bench_tilemodel.err
timeout 300 $B > $O/ab_bench_default2.json 2> $O/ab_bench_default2.err
timeout 400 $B --workload yelp_large > $O/ab_bench_large_fused.json 2> $O/ab_bench_large_fused.err
P5_NO_FATTN_BWD_BIG=1 timeout 400 $B --workload yelp_large > $O/ab_bench_large_mat.json 2> $O/ab_bench_large_mat.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        ev = d.get("eval") or {}
        print("%-34s train %8.1f samples/s %7.3f ms/step  gemm256 %.0f TF/s  eval %s" % (
            f.split("/")[-1], d["value"], d["ms_per_step"], d["roofline"]["achieved"],
            ("%.0f items/s %.3f ms/batch" % (ev["value"], ev["ms_per_batch"])) if ev else "-"))
    except Exception as e:
        print(f, "failed", repr(e)[:200])
PY
