"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (all kernels in the report).
usage: python scripts/ncu_hot_sass.py report.ncu-rep [kernel-substring] [topN]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    blocks, cur = [], None
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = {"name": line, "lines": []}
            blocks.append(cur)
        elif cur is not None:
            cur["lines"].append(line)
    for b in blocks:
        if sub not in b["name"]:
            continue
        rows = list(csv.reader(io.StringIO("\n".join(b["lines"]))))
        hdr, rows = rows[0], rows[1:]
        i_src, i_s = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
        tot = sum(int(r[i_s]) for r in rows)
        print("==", b["name"][:140], "total samples", tot)
        order = sorted(range(len(rows)), key=lambda i: -int(rows[i][i_s]))[:top]
        for i in order:
            r = rows[i]
            reasons = sorted(((int(r[c]), hdr[c][6:]) for c in stall_cols if r[c] not in ("", "0")), reverse=True)[:3]
            print("%6.2f%%  #%-5d %-60s %s" % (100.0 * int(r[i_s]) / max(tot, 1), i, r[i_src].strip()[:60],
                                           " ".join("%s:%d" % (n, v) for v, n in reasons)))


if __name__ == "__main__":
    main()
