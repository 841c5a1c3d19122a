"""Top stall-sample instructions of one kernel instance in an .ncu-rep:  python tools/ncu_source_top.py <rep> <kernel-id filter> [N]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", kid], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
tables, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        tables.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
t = tables[0]
hdr = t["rows"][0]
ia, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = [(int(r[isamp] or 0), r[ia].strip(), int(r[iex] or 0), k) for k, r in enumerate(t["rows"][1:]) if len(r) > max(isamp, iex)]
tot = sum(d[0] for d in data) or 1
print(t["name"][:100]); print("total samples", tot, "instructions", sum(d[2] for d in data), "tables", len(tables))
for d in sorted(data, reverse=True)[:n]:
    print("%5.1f%% %9d  #%-5d %s" % (100.0 * d[0] / tot, d[2], d[3], d[1][:100]))
