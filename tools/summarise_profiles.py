"""Turn the artefacts of tools/collect_profiles.sh (gpurun_out/) into the tracked summaries under profiles/.
usage: python tools/summarise_profiles.py [tag]      (needs `ncu` for --page raw export of the .ncu-rep files)"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.environ.get("P5_PROF_DIR", os.path.join(ROOT, "profiles"))     # on the GPU box: a directory under gpurun_out/ (only that comes back)
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def launches(path, dst, title):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        n = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("p5::", "").replace("<unnamed>::", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("%s\nsum of kernel durations %.1f us over %d launches (ncu serialises kernels: shares, not absolutes)\n" %
                (title, tot, sum(v[0] for v in agg.values())))
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%10.1f us %5.1f%% %5d x %8.2f us  %s\n" % (t, 100 * t / tot, c, t / c, n[:90]))


def ncu_summary(rep, dst):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return False
    h, u = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("source: %s  (ncu --set full --clock-control none; cold-cache, serialised launch)\n" % os.path.basename(rep))
        for v in rows[2:]:
            d = dict(zip(h, v))
            f.write("---\nKernel Name = %s\n" % d.get("Kernel Name", "")[:160])
            for m in METRICS:
                if m in d:
                    f.write("%s = %s %s\n" % (m, d[m], u[h.index(m)]))
    return True


def gemm_traffic(rep, dst):
    """DRAM bytes of the first captured gemm_tc_kernel<256, .> launch -> profiles/<tag>_gemm_traffic.json (bench.py reads it)."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        d = dict(zip(h, v))
        if "gemm_tc_kernel<256" not in d.get("Kernel Name", ""):
            continue

        def to_bytes(name):
            val = float(d[name].replace(",", ""))
            unit = u[h.index(name)].lower()
            return val * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        out = {"kernel": d["Kernel Name"][:120], "launch": "grid %s block %s" % (d.get("Grid Size"), d.get("Block Size")),
               "dram_bytes": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
               "duration_us_under_ncu": float(d["gpu__time_duration.sum"].replace(",", ""))}
        json.dump(out, open(dst, "w"), indent=1)
        return


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    b = os.path.join(OUT, tag + "_bench_n1.log")
    if os.path.exists(b):
        line = [l for l in open(b).read().splitlines() if l.startswith("{")][-1]
        json.dump(json.loads(line), open(os.path.join(PROF, tag + "_bench_n1.json"), "w"), indent=1)
    for kind, title in (("train_step", "one T5-base train step (B=64, Le=256 packed, bf16, dropout 0.1)"),
                        ("eval_batch", "one eval batch (20 users x 20 beams, T5-base, trie-constrained)")):
        src = os.path.join(OUT, "%s_%s_launches.csv" % (tag, kind))
        if os.path.exists(src):
            shutil.copy(src, os.path.join(PROF, os.path.basename(src)))
            launches(src, os.path.join(PROF, "%s_%s_kernel_shares.txt" % (tag, kind)), title)
    for name in ("gemm", "fattn_fwd", "fattn_bwd", "dattn_fwd", "dattn_bwd", "rmsnorm_fwd", "rmsnorm_bwd", "adamw", "ce_fwd", "ce_bwd",
                 "embed_fwd", "embed_bwd", "sumsq", "decode_persistent", "gemm_qkv", "gemm_wgrad"):
        rep = os.path.join(OUT, "%s_%s.ncu-rep" % (tag, name))
        if os.path.exists(rep):
            ncu_summary(rep, os.path.join(PROF, "%s_%s_ncu_full_summary.txt" % (tag, name)))
            if name == "gemm":
                gemm_traffic(rep, os.path.join(PROF, "%s_gemm_traffic.json" % tag))


if __name__ == "__main__":
    main()
