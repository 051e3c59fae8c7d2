#!/bin/bash
# Final round-2 evidence on ONE B200 (under gpurun): the whole GPU test suite, the bench (both arms), the configs[2]/[3]
# workloads, launch lists of a train step / an eval batch and `ncu --set full` captures of the kernels that changed in
# this session.  Everything lands in gpurun_out/ (summaries: gpurun_out/profiles_r02/).
set -u
mkdir -p gpurun_out
O=gpurun_out
TAG=r02
timeout 900 python -m pytest tests -q -m gpu -x > $O/${TAG}_pytest_gpu.log 2>&1
tail -3 $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.log 2> $O/${TAG}_bench_n1.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --workload yelp_large > $O/${TAG}_bench_yelp_large.json 2> $O/${TAG}_bench_yelp_large.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --workload beauty_collab > $O/${TAG}_bench_beauty_collab.json 2> $O/${TAG}_bench_beauty_collab.err
NCU="ncu --profile-from-start off --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_train_step_launches.csv python tools/profile_step.py > $O/${TAG}_ncu_train.log 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_eval_batch_launches.csv python tools/profile_step.py eval > $O/${TAG}_ncu_eval.log 2>&1
for K in "gemm_tc_kernel:gemm" "fattn_fwd_kernel:fattn_fwd" "fattn_bwd_kernel:fattn_bwd" "dattn_fwd_kernel:dattn_fwd" "dattn_bwd_kernel:dattn_bwd" "rmsnorm_bwd:rmsnorm_bwd" "sumsq_partial_kernel:sumsq"; do
  PAT=${K%%:*}; NAME=${K##*:}
  timeout 240 $NCU --set full -k regex:$PAT -c 1 -f -o $O/${TAG}_${NAME} python tools/profile_step.py > $O/${TAG}_ncu_${NAME}.log 2>&1
done
timeout 400 $NCU --set full -k regex:decode_persistent_kernel -c 1 -f -o $O/${TAG}_decode_persistent python tools/profile_step.py eval > $O/${TAG}_ncu_decode.log 2>&1
P5_PROF_DIR=$O/profiles_${TAG} python tools/summarise_profiles.py ${TAG}
rm -f $O/${TAG}_*.ncu-rep
timeout 200 python tools/decode_phases.py > $O/${TAG}_decode_phases.log 2>&1
timeout 200 python tools/gemm_shapes_in_step.py > $O/${TAG}_gemm_shapes_in_step.log 2>&1
python - <<'PY'
import json, glob
for f in ["gpurun_out/r02_bench_n1.log", "gpurun_out/r02_bench_ref.json", "gpurun_out/r02_bench_yelp_large.json", "gpurun_out/r02_bench_beauty_collab.json"]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        ev = d.get("eval") or {}
        print("%-40s %10.1f %s %8.3f ms/step  e2e %s  roofline %s  eval %s" % (
            f.split("/")[-1], d["value"], d["unit"], d.get("ms_per_step", 0), (d.get("e2e") or {}).get("value"),
            (d.get("roofline") or {}).get("frac"), ("%.0f items/s %.3f ms/batch frac %s" % (ev["value"], ev["ms_per_batch"], (ev.get("roofline") or {}).get("frac"))) if ev else "-"))
    except Exception as e:
        print(f, "failed", repr(e)[:200])
PY
du -sh $O
