#!/bin/bash
# 2-GPU evidence (gpurun --gpus 2): the NCCL gradient parity test, then the weak-scaling pair N=1 / N=2 on the SAME box and
# an A/B of the CTAs handed to NCCL (P5_COMM_CTAS=8: 8 CTAs for the all-reduce and 8 SMs kept free of persistent GEMM CTAs).
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests -m gpu -q -k "two_rank" 2>&1 | tail -4 > gpurun_out/r02_n2_pytest.log
cat gpurun_out/r02_n2_pytest.log
timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_scale_n2.json 2> gpurun_out/r02_scale_n2.err
P5_COMM_CTAS=8 timeout 600 $TR --master-port 29513 bench.py --gpus 2 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_scale_n2_ctas8.json 2> gpurun_out/r02_scale_n2_ctas8.err
python - <<'PY'
import json
for n in ("n1", "n2", "n2_ctas8"):
    try:
        d = json.loads([l for l in open("gpurun_out/r02_scale_%s.json" % n) if l.startswith("{")][-1])
        print(n, "train %.1f samples/s (%.3f ms/step)  eval %.0f items/s" % (d["value"], d["ms_per_step"], d["eval"]["value"]))
    except Exception as e:
        print(n, "failed", e)
PY
