#!/bin/bash
# 2-GPU check of the final round-2 code (gpurun --gpus 2): the NCCL gradient parity test (fused train step: all-reduce per
# range + early gradient norm on the communication stream), then N = 1 and N = 2 on the same box
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests -m gpu -q -k "two_rank" 2>&1 | tail -4 > gpurun_out/r02f_n2_pytest.log
cat gpurun_out/r02f_n2_pytest.log
timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02f_scale_n1.json 2> gpurun_out/r02f_scale_n1.err
timeout 300 $TR --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02f_scale_n2.json 2> gpurun_out/r02f_scale_n2.err
python - <<'PY'
import json
for n in ("n1", "n2"):
    try:
        d = json.loads([l for l in open("gpurun_out/r02f_scale_%s.json" % n) if l.startswith("{")][-1])
        print(n, "train %.1f samples/s (%.3f ms/step)  eval %.0f items/s  replica diff %s" % (d["value"], d["ms_per_step"], d["eval"]["value"], d.get("dp_replica_max_rel_diff")))
    except Exception as e:
        print(n, "failed", repr(e)[:300])
PY
tail -3 gpurun_out/r02f_scale_n2.err
