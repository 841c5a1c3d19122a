"""profiles/<tag>_launches.md and profiles/<tag>_gemm_per_launch.md from an ncu launch list of whole fused cfg2 steps.

Input: gpurun_out/launches_<tag>.csv = `GANTTS_B200_CUDA_PROFILE_STEPS=1 ncu --profile-from-start off --metrics
gpu__time_duration.sum --clock-control none --csv --log-file ... python bench.py --steps 2 --warmup 3 --no-cpu-baseline
--no-dropin` (exactly one step, launch order).  Optional: gpurun_out/prof_step_<tag>.ncu-rep (`--set full` of the same
window) for the tensor-pipe / DRAM columns.  Usage: python tools/per_launch.py r02 [bench-json-for-in-step-times]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import launch_summary  # noqa: E402

TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
GO, OUT = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
M = 32000
PEAK_TF = 1425.5
HBM_TBS = 6.48
try:
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    PEAK_TF = pk.get("bf16_tflops_sustained", PEAK_TF)
    HBM_TBS = pk.get("hbm_gbs", HBM_TBS * 1e3) / 1e3
except Exception:
    pass

# K-major launches (y = x W^T or gx = gz W) and MN-major launches (gW = gz^T x) of one step, in launch order per kind:
# (label, rows, N, K, output kind)
KK = [("G fwd 425->512 (+act, planes)", M, 512, 425, "planes"), ("G fwd 512->512", M, 512, 512, "planes"),
      ("G fwd 512->512", M, 512, 512, "planes"), ("G fwd 512->187 (fp32 y_hat, ld 187)", M, 187, 512, "f32"),
      ("D fwd 58->256, real|fake 2M rows", 2 * M, 256, 58, "planes"), ("D fwd 256->256, 2M rows", 2 * M, 256, 256, "planes"),
      ("D fwd 256->256, 2M rows", 2 * M, 256, 256, "planes"),
      ("D bwd gx3 = gz W", 2 * M, 256, 256, "planes"), ("D bwd gx2", 2 * M, 256, 256, "planes"),
      ("D bwd gx1 (fake half, fp32, ld 58)", M, 58, 256, "f32"),
      ("D(adv) fwd 58->256, M rows", M, 256, 58, "planes"), ("D(adv) fwd 256->256", M, 256, 256, "planes"),
      ("D(adv) fwd 256->256", M, 256, 256, "planes"),
      ("D(adv) bwd gx3", M, 256, 256, "planes"), ("D(adv) bwd gx2", M, 256, 256, "planes"),
      ("D(adv) bwd gx1 -> += g_static window (fp32)", M, 58, 256, "f32"),
      ("G bwd gx4 = gz W (512 wide)", M, 512, 187, "planes"), ("G bwd gx3", M, 512, 512, "planes"),
      ("G bwd gx2", M, 512, 512, "planes")]
MN = [("D bwd gW3 = gz^T h (256x256)", 2 * M, 256, 256), ("D bwd gW2", 2 * M, 256, 256), ("D bwd gW1 (256x58)", 2 * M, 256, 58),
      ("G bwd gW4 (187x512)", M, 187, 512), ("G bwd gW3 (512x512)", M, 512, 512), ("G bwd gW2 (512x512)", M, 512, 512),
      ("G bwd gW1 (512x425)", M, 512, 425)]


def short(name):
    n = name.replace("gantts::", "").replace("void ", "")
    cut = n.rfind("(")
    return (n[:cut] if cut > 0 else n).replace("(bool)", "").replace("(int)", "")[:64]


def full_metrics():
    rep = os.path.join(GO, "prof_step_%s.ncu-rep" % TAG)
    raw = os.path.join(GO, "step_full_%s.csv" % TAG)
    if os.path.exists(raw):
        out = open(raw).read()
    elif os.path.exists(rep):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return None
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return None
    h, units = rows[0], rows[1]
    want = {"tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "inst": "smsp__inst_executed.sum", "us": "gpu__time_duration.sum", "name": "Kernel Name",
            "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed"}
    idx = {k: h.index(v) for k, v in want.items() if v in h}
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    recs = []
    for r in rows[2:]:
        d = {}
        for k, i in idx.items():
            d[k] = r[i]
        for k in ("rd", "wr"):
            if k in idx:
                d[k] = float(d[k].replace(",", "")) * scale.get(units[idx[k]], 1.0)
        recs.append(d)
    return recs


def main():
    seq = launch_summary.load(os.path.join(GO, "launches_%s.csv" % TAG))
    full = full_metrics()
    if full is not None and len(full) != len(seq):
        print("note: --set full capture has %d launches, the launch list %d; full columns dropped" % (len(full), len(seq)))
        full = None
    tot = sum(v for _, _, v in seq)
    lines = ["# ncu launch list of ONE fused cfg2 step, %s (`ncu --profile-from-start off --metrics gpu__time_duration.sum "
             "--clock-control none`, window = bench.py's GANTTS_B200_CUDA_PROFILE_STEPS=1)" % TAG, "",
             "%d launches, %.1f us summed (cold-cache, serialised: compare SHARES with the in-step CUDA-event times of "
             "profiles/%s_bench.json, not absolutes)." % (len(seq), tot, TAG), "",
             "| # | kernel | grid | us | share |" + (" tensor pipe % | issue % | DRAM rd MB | DRAM wr MB | DRAM GB/s |" if full else ""),
             "|---|---|---|---|---|" + ("---|---|---|---|---|" if full else "")]
    agg = {}
    for i, (name, grid, v) in enumerate(seq):
        extra = ""
        if full:
            f = full[i]
            extra = " %.1f | %.1f | %.1f | %.1f | %.0f |" % (float(f.get("tensor", 0) or 0), float(f.get("issue", 0) or 0), f["rd"], f["wr"],
                                                          (f["rd"] + f["wr"]) * 1e-3 / (float(f["us"]) * 1e-6) if float(f["us"]) else 0)
        lines.append("| %d | `%s` | %s | %.1f | %.3f |%s" % (i, short(name), grid.replace(", 1, 1)", ")"), v, v / tot, extra))
        a = agg.setdefault(short(name), [0, 0.0])
        a[0] += 1
        a[1] += v
    lines += ["", "## per kernel", "", "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.1f | %.1f | %.3f |" % (k, n, t, t / n, t / tot))
    open(os.path.join(OUT, "%s_launches.md" % TAG), "w").write("\n".join(lines) + "\n")

    # ---- GEMM per-launch roofline view
    kk, mn = list(KK), list(MN)
    rows, gsum, fl_sum = [], 0.0, 0.0
    for i, (name, grid, v) in enumerate(seq):
        n = short(name)
        is_mn = "gemm_pair_mn_kernel" in n or "gemm_bf16x3_kernel<true" in n or "gemm_bf16x3_kernel<1" in n
        is_kk = not is_mn and ("gemm_pair_kernel" in n or "gemm_bf16x3_kernel" in n)
        if not (is_mn or is_kk):
            continue
        if is_mn and mn:
            label, rws, N, K = mn.pop(0)
            out_bytes = N * K * 4.0
            in_bytes = rws * (N + K) * 4.0
        elif is_kk and kk:
            label, rws, N, K, okind = kk.pop(0)
            in_bytes = rws * K * 4.0 + N * K * 4.0
            out_bytes = rws * N * 4.0 + (rws * N / 4.0 if okind == "planes" else 0.0)
        else:
            label, rws, N, K, in_bytes, out_bytes = "(unlabelled)", 0, 0, 0, 0.0, 0.0
        flops = 3.0 * 2.0 * rws * N * K
        tf = flops / (v * 1e-6) / 1e12 if v else 0.0
        hbm_floor = (in_bytes + out_bytes) / (HBM_TBS * 1e12) * 1e6
        tc_floor = flops / (PEAK_TF * 1e12) * 1e6
        extra = ""
        if full:
            f = full[i]
            extra = " %.1f | %.0f |" % (float(f.get("tensor", 0) or 0), f["rd"] + f["wr"])
        rows.append("| %d | %s | %s | %d x %d x %d | %.1f | %.0f | %.0f %% | %.1f | %.1f | %s |%s" % (
            i, label, "MN-major" if is_mn else ("pair" if "pair" in n else "K-major"), rws, N, K, v, tf, 100.0 * tf / PEAK_TF,
            hbm_floor, tc_floor, "tensor" if tc_floor > hbm_floor else "hbm", extra))
        gsum += v
        fl_sum += flops
    hdr = ["# Per-launch view of the tcgen05 GEMM launches of ONE fused cfg2 step (%s)" % TAG, "",
           "Times: `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: upper bounds of the in-step "
           "times). Executed flops = 3 x 2MNK (bf16x3). `HBM floor` = unique operand + output bytes (4 B per element for hi/lo planes "
           "and fp32, + 2 bits per element of derivative codes) / %.2f TB/s; `tensor floor` = executed flops / %.1f TF/s "
           "(MEASURED_PEAKS.json, sustained bf16)." % (HBM_TBS, PEAK_TF), "",
           "| # | launch | kernel | rows x N x K | us | executed TF/s | % of sustained bf16 peak | HBM floor us | tensor floor us | bound |"
           + (" tensor pipe active % | DRAM MB |" if full else ""),
           "|---|---|---|---|---|---|---|---|---|---|" + ("---|---|" if full else "")]
    tail = ["", "%d GEMM launches, %.1f us summed = %.0f TF/s executed = %.3f of the sustained bf16 peak (in-step, by CUDA events: "
            "profiles/%s_bench.json roofline.gemm_family)." % (len(rows), gsum, fl_sum / (gsum * 1e-6) / 1e12 if gsum else 0,
                                                               fl_sum / (gsum * 1e-6) / 1e12 / PEAK_TF if gsum else 0, TAG)]
    if kk or mn:
        tail.append("unmatched labels: %d K-major, %d MN-major (launch structure changed; labels above may be shifted)" % (len(kk), len(mn)))
    open(os.path.join(OUT, "%s_gemm_per_launch.md" % TAG), "w").write("\n".join(hdr + rows + tail) + "\n")
    print("\n".join(hdr[-2:] + rows + tail))


if __name__ == "__main__":
    main()
