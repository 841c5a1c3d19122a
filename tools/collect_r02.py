"""gpurun_out/ scratch of tools/gpu_session.sh  ->  the committed round-2 evidence under profiles/.
python tools/collect_r02.py [tag]   (after the session's files have been merged back into gpurun_out/)"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
GO, OUT = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def last_json(path):
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def copy_bench():
    names = {"": "", "_reference": "_reference", "_cfg3": "_cfg3", "_cfg5": "_cfg5", "_n2": "_n2", "_n4": "_n4", "_n8": "_n8"}
    got = {}
    for src, dst in names.items():
        d = last_json(os.path.join(GO, "bench_%s%s.json" % (TAG, src)))
        if d is not None:
            json.dump(d, open(os.path.join(OUT, "%s_bench%s.json" % (TAG, dst)), "w"), indent=1)
            got[src] = d
    return got


def top_kernel():
    raw = os.path.join(GO, "step_full_%s.csv" % TAG)
    if not os.path.exists(raw):
        return
    rows = list(csv.reader(open(raw).read().splitlines()))
    h, u = rows[0], rows[1]
    i_name, i_rd, i_wr, i_us = (h.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                                      "gpu__time_duration.sum"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n, us = 0.0, 0, 0.0
    for r in rows[2:]:
        nm = r[i_name]
        kmaj = ("gemm_pair_kernel" in nm or "gemm_bf16x3_kernel<(bool)0" in nm or "gemm_bf16x3_kernel<false" in nm or
                "gemm_bf16x3_kernel<0," in nm)
        if not kmaj:
            continue
        tot += float(r[i_rd].replace(",", "")) * scale.get(u[i_rd], 1.0) + float(r[i_wr].replace(",", "")) * scale.get(u[i_wr], 1.0)
        us += float(r[i_us].replace(",", ""))
        n += 1
    if n:
        json.dump({"kernel": "K-major tcgen05 GEMM launches of one fused cfg2 step (gemm_pair_kernel / gemm_bf16x3_kernel<K-major>)",
                   "launches": n, "dram_bytes_per_launch": tot / n, "avg_duration_us_under_ncu": us / n,
                   "source": "gpurun_out/step_full_%s.csv = ncu --set full --clock-control none over the profiler window of one "
                             "step (tools/gpu_session.sh); per-launch rows in profiles/%s_launches.md" % (TAG, TAG)},
                  open(os.path.join(OUT, "%s_top_kernel.json" % TAG), "w"), indent=1)


def scaling(b):
    if "" not in b or not any(k in b for k in ("_n2", "_n4", "_n8")):
        return
    base = b[""]
    # the N > 1 runs use --steps 50 --warmup 10; the N = 1 line with the same arguments is kept as
    # <tag>_bench_scaling_n1.json (the default 100-step run of <tag>_bench.json reaches the software power cap)
    sb = os.path.join(OUT, "%s_bench_scaling_n1.json" % TAG)
    base_file = "%s_bench.json" % TAG
    if os.path.exists(sb):
        base = json.load(open(sb))
        base_file = "%s_bench_scaling_n1.json (--steps 50 --warmup 10, like the rows below)" % TAG
    lines = ["# Weak scaling on one B200 node, %s (`bench.py --gpus N` under torchrun, one rank per GPU, B=32 x T=1000 per GPU)" % TAG, "",
             "value = total frames of all ranks / max-over-ranks CUDA-event time, median of >= 3 repeats after 12 untimed all-reduces of "
             "both gradient buffers.  Exchange per step: SUM all-reduce of the D buffer (0.6 MB) after phase 1 and of the G buffer "
             "(3.4 MB) after phase 2 of `gantts_gan_step`; `allreduce` = the two all-reduces timed alone, back to back.", "",
             "| N | ms/step | frames/s (device-resident) | vs N=1 (efficiency) | e2e ms/step | all-reduce pair alone us | file |",
             "|---|---|---|---|---|---|---|",
             "| 1 | %.4f | %.2f M | 1.00 | %.3f | - | %s |" % (base["ms_per_step"], base["value"] / 1e6, base["e2e"]["ms_per_step"], base_file)]
    for k, n in (("_n2", 2), ("_n4", 4), ("_n8", 8)):
        if k in b:
            d = b[k]
            ar = d.get("allreduce") or {}
            lines.append("| %d | %.4f | %.2f M | %.2f (%.0f %%) | %.3f | %s | %s_bench%s.json |" % (
                n, d["ms_per_step"], d["value"] / 1e6, d["value"] / base["value"], 100.0 * d["value"] / base["value"] / n,
                d["e2e"]["ms_per_step"], ("%.0f" % ar["us_per_step_pair"]) if ar else "-", TAG, k))
    c5, c5n = last_json(os.path.join(GO, "bench_%s_cfg5.json" % TAG)), last_json(os.path.join(GO, "bench_%s_cfg5_n8.json" % TAG))
    if c5 and c5n:
        json.dump(c5n, open(os.path.join(OUT, "%s_bench_cfg5_n8.json" % TAG), "w"), indent=1)
        sb5 = os.path.join(OUT, "%s_bench_cfg5_scaling_n1.json" % TAG)
        note5 = ""
        if os.path.exists(sb5):
            cur5, c5 = c5, json.load(open(sb5))
            note5 = "  (Both on build c21373d; the final build runs cfg5 at %.1f ms/step on one GPU.)" % cur5["ms_per_step"]
        lines += ["", "cfg5 (BASELINE configs[4]: TTS acoustic LSTMRNN + GAN + MGE, B=64 x T=1500 per GPU, `GanTrainer`, gradient all-reduce of "
                  "the 17 M-parameter generator and the discriminator per step): 1 GPU %.1f ms/step = %.1f k frames/s (%s_bench_cfg5*.json); "
                  "**8 GPUs %.1f ms/step = %.2f M frames/s = %.2fx (%.0f %%)** (%s_bench_cfg5_n8.json)." % (
                      c5["ms_per_step"], c5["value"] / 1e3, TAG, c5n["ms_per_step"], c5n["value"] / 1e6, c5n["value"] / c5["value"],
                      100.0 * c5n["value"] / c5["value"] / 8, TAG) + note5]
    open(os.path.join(OUT, "%s_scaling.md" % TAG), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    b = copy_bench()
    top_kernel()
    scaling(b)
    if os.path.exists(os.path.join(GO, "launches_%s.csv" % TAG)):
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "per_launch.py"), TAG], check=False, stdout=subprocess.DEVNULL)
    for f in ("tests_%s.log" % TAG, "smoke_%s.log" % TAG):
        p = os.path.join(GO, f)
        if os.path.exists(p):
            open(os.path.join(OUT, f.replace(".log", ".txt")), "w").write(open(p).read())
    print(sorted(f for f in os.listdir(OUT) if f.startswith(TAG)))
