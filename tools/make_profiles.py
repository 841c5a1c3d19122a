"""Turn the scratch outputs of tools/gpu_session.sh (gpurun_out/) into the committed summaries under profiles/."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import launch_summary  # noqa: E402

TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")


def launches():
    seq = launch_summary.load(os.path.join(GO, "launches_%s.csv" % TAG))
    agg = {}
    for name, grid, v in seq:
        a = agg.setdefault(name.split("(")[0][:80], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = ["# ncu launch list, %s (`ncu --metrics gpu__time_duration.sum --clock-control none`, one fused GAN step ~ "
             "%d launches; cold-cache serialised times: compare SHARES)" % (TAG, len(seq)), "",
             "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.1f | %.1f | %.3f |" % (k, n, t, t / n, t / tot))
    lines.append("")
    lines.append("total %.1f us over %d launches" % (tot, len(seq)))
    open(os.path.join(OUT, "%s_launches.md" % TAG), "w").write("\n".join(lines) + "\n")
    return agg, tot


def gemm_full():
    rep = os.path.join(GO, "prof_gemm_%s.ncu-rep" % TAG)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    idx = [hdr.index(w) for w in want]
    units = rows[1]
    recs = []
    for r in rows[2:]:
        rec = {w: r[i] for w, i in zip(want, idx)}
        rec["_units"] = {w: units[i] for w, i in zip(want, idx)}
        recs.append(rec)
    lines = ["# ncu --set full, tcgen05 GEMM launches of one fused step (%s)" % TAG, "",
             "| # | kernel<MN,EPI> | us | tensor pipe active % | dram rd MB | dram wr MB | L2->SM rd MB | warp inst | issue active % |",
             "|---|---|---|---|---|---|---|---|---|"]

    def mb(v, u):
        f = float(v)
        return f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    best = None
    for n, r in enumerate(recs):
        name = r["Kernel Name"]
        tmpl = name[name.find("<"):name.find(">") + 1]
        dr = mb(r["dram__bytes_read.sum"], r["_units"]["dram__bytes_read.sum"])
        dw = mb(r["dram__bytes_write.sum"], r["_units"]["dram__bytes_write.sum"])
        l2 = float(r["lts__t_sectors_srcunit_tex_op_read.sum"]) * 32e-6
        us = float(r["gpu__time_duration.sum"])
        lines.append("| %d | %s | %.1f | %.1f | %.1f | %.1f | %.1f | %s | %.1f |" % (
            n, tmpl, us, float(r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]), dr, dw, l2,
            r["smsp__inst_executed.sum"].split(".")[0], float(r["smsp__issue_active.avg.pct_of_peak_sustained_active"])))
        if "(bool)0, (int)1" in name or "<0, 1" in name:
            if best is None or us > best[0]:
                best = (us, dr + dw, r)
    open(os.path.join(OUT, "%s_gemm_ncu_full.md" % TAG), "w").write("\n".join(lines) + "\n")
    if best:
        json.dump({"kernel": "gemm_bf16x3_kernel<K-major, EPI_PLANES_FWD> (generator hidden layer, M=32000 N=512 K=512)",
                   "duration_us_under_ncu": best[0], "dram_bytes_per_launch": best[1] * 1e6,
                   "algorithmic_bytes_per_launch": 32000 * 512 * 4.0 * 2 + 512 * 512 * 4.0,
                   "source": "profiles/%s_gemm_ncu_full.md (ncu --set full --clock-control none)" % TAG},
                  open(os.path.join(OUT, "%s_top_kernel.json" % TAG), "w"), indent=1)


def bench():
    for suffix in ("", "_reference"):
        p = os.path.join(GO, "bench_%s%s.json" % (TAG, suffix))
        if os.path.exists(p):
            lines = [l for l in open(p).read().splitlines() if l.startswith("{")]
            if lines:
                json.dump(json.loads(lines[-1]), open(os.path.join(OUT, "%s_bench%s.json" % (TAG, suffix)), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    launches()
    gemm_full()
    bench()
    print(os.listdir(OUT))
