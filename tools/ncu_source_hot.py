"""Top stall-sample instructions per kernel launch from `ncu -i rep --page source --csv` (SASS view)."""
import csv
import subprocess
import sys


def main():
    rep, which = sys.argv[1], int(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(out.splitlines()):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "hdr": None}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None:
            cur["rows"].append(row)
    b = blocks[which]
    h = b["hdr"]
    i_src, i_samp, i_exec = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_")]
    print(b["name"][:100], "instructions:", len(b["rows"]))
    tot = sum(int(r[i_samp] or 0) for r in b["rows"])
    print("total samples", tot)
    rows = sorted(enumerate(b["rows"]), key=lambda t: -int(t[1][i_samp] or 0))[:top]
    for idx, r in rows:
        st = sorted(((int(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
        print("%5d %6s %7s  %-60s %s" % (idx, r[i_samp], r[i_exec], r[i_src].strip()[:60], st))


if __name__ == "__main__":
    main()
