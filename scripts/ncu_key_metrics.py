"""Key metrics of an `ncu --page raw --csv` export, one block per kernel launch.
usage: python scripts/ncu_key_metrics.py raw.csv [--json]"""
import csv
import json
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__cycles_active.avg", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for d in data:
        rec = {"kernel": d[hdr.index("Kernel Name")]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                rec[w] = "%s %s" % (d[i], units[i])
        out.append(rec)
    if "--json" in sys.argv:
        print(json.dumps(out, indent=1))
    else:
        for rec in out:
            print("== " + rec["kernel"])
            for k, v in rec.items():
                if k != "kernel":
                    print("  %-80s %s" % (k, v))


if __name__ == "__main__":
    main()
