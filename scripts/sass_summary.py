"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel, from the built product
library: UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk), UTMALDG / UTMASTG (tensor-map TMA),
UTCBAR (tcgen05.commit), HMMA (legacy mma.sync), FFMA (fp32 SIMT math).  Runs here (no GPU): cuobjdump -sass.
usage: python scripts/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "FFMA"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "melgan_multi_b200", "lib", "libmelgan_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[cur][mn] += 1
    names = subprocess.run(["c++filt"], input="\n".join(order), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    print("# %s  (cuobjdump -sass; static instruction counts per kernel)" % os.path.relpath(lib, ROOT))
    print("# %-118s %s" % ("kernel", " ".join("%7s" % m for m in MNEMONICS)))
    tot = collections.Counter()
    for mangled, name in sorted(zip(order, names), key=lambda t: t[1]):
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\(.*$", "", name)
        c = counts[mangled]
        tot.update(c)
        print("%-120s %s" % (name[:120], " ".join("%7d" % c[m] for m in MNEMONICS)))
    print("%-120s %s" % ("TOTAL", " ".join("%7d" % tot[m] for m in MNEMONICS)))


if __name__ == "__main__":
    main()
