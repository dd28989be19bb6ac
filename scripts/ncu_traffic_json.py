"""DRAM traffic per launch of the generator's chain kernels from an `ncu --set full ... --page raw --csv` export of
scripts/one_forward_each.py (single chain: MG_GEN_SLICES=1), as JSON keyed by the chain's kernel names, each entry tagged
with the template configuration the library reports for that kernel (mg_gen_kernel_config): bench.py quotes an entry as
`roofline.traffic` only if the configuration still matches the build it runs.  Run ON THE GPU BOX right after the capture.
usage: python scripts/ncu_traffic_json.py raw.csv > profiles/rNN_ncu_traffic.json"""
import csv
import ctypes
import json
import sys

sys.path.insert(0, ".")
from melgan_multi_b200 import engine


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}

    def val(d, name, scale):
        v, u = float(d[col[name]]), units[col[name]]
        return v * scale.get(u, 1.0)
    L = engine.lib()
    L.mg_gen_kernel_config.restype = ctypes.c_char_p
    L.mg_gen_kernel_config.argtypes = [ctypes.c_int, ctypes.c_int]
    n = L.mg_gen_forward_launches()
    gen = [d for d in data if any(k in d[col["Kernel Name"]] for k in ("resblock_tc_kernel", "convt_tc_kernel", "convt_resident_tc_kernel"))
           or "ConvCfg<80" in d[col["Kernel Name"]]][:n]
    out = {"_comment": "dram__bytes_read.sum / dram__bytes_write.sum per launch, ncu --set full --clock-control none, one generator "
                       "forward at config 2 (B=64, T=32) as ONE chain (MG_GEN_SLICES=1); kernel_config = mg_gen_kernel_config(i, 32) "
                       "of the build that was profiled"}
    for i, d in enumerate(gen):
        out[L.mg_gen_kernel_name(i).decode()] = {
            "kernel": d[col["Kernel Name"]][:160], "kernel_config": L.mg_gen_kernel_config(i, 32).decode(),
            "dram_read_bytes": val(d, "dram__bytes_read.sum", {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}),
            "dram_write_bytes": val(d, "dram__bytes_write.sum", {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}),
            "gpu_time_us": val(d, "gpu__time_duration.sum", {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}),
            "tensor_pipe_active_pct_of_active": float(d[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
        }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
