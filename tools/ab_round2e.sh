#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
CASES="bwd_bf16_small bwd_bf16_base_le256_packed dropout_bf16_small xcheck_dattn_dropout_small xcheck_fbwd_dropout_base_le256_packed bwd_bf16_c2full_packed bwd_fp32_small adamw_fp32_tiny bwd_bf16_large_le128"
timeout 600 python tests/gpu_cases_model.py $CASES > $O/ab4_cases.log 2>&1
tail -1 $O/ab4_cases.log
grep -v '"ok": true' $O/ab4_cases.log | cut -c1-600 | head -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > $O/ab4_bench.json 2> $O/ab4_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/ab4_bench.json") if l.startswith("{")][-1])
print("train %.1f samples/s %.3f ms/step gemm256 %.0f TF/s eval %.0f items/s %.3f ms/batch" % (d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["eval"]["value"], d["eval"]["ms_per_batch"]))
PY
