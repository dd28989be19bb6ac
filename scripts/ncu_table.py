"""One line per kernel launch from an `ncu --page raw --csv` export: time, DRAM bytes, tensor-pipe / SM / L2 utilisation,
occupancy, registers, grid.  Usage: python scripts/ncu_table.py raw.csv "title line" > profiles/....txt"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB_rd"), ("dram__bytes_write.sum", "MB_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid")]
FMT = "%-78s %8s %8s %8s %6s %10s %6s %6s %6s %5s %6s"


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    print("# " + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
    print("# per-launch times are cold-cache and serialised (ncu replays): compare shares, not absolutes")
    print(FMT % tuple(["kernel"] + [c[1] for c in COLS]))
    for d in data:
        name = d[col["Kernel Name"]].replace("void ", "").replace("mg::", "")[:78]
        vals = []
        for c, _ in COLS:
            v, u = d[col[c]], units[col[c]]
            try:
                f = float(v)
                if c.startswith("dram__bytes"):
                    f = {"Mbyte": f, "Kbyte": f / 1e3, "Gbyte": f * 1e3}.get(u, f / 1e6)
                if c == "gpu__time_duration.sum" and u == "ms":
                    f *= 1e3
                vals.append("%.1f" % f)
            except ValueError:
                vals.append(v)
        print(FMT % tuple([name] + vals))


if __name__ == "__main__":
    main()
