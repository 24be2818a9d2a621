"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum + dram bytes per launch) per kernel family.

    python tools/ncu_launch_summary.py gpurun_out/launches.csv [out.json]
"""
import collections
import csv
import json
import re
import sys


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    per = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"[<(].*", "", r["Kernel Name"]).replace("void ", "").replace("st::", "")
        if not name.startswith(("gemm", "attn", "gn_", "layernorm", "geglu", "linear_small", "conv3x3", "im2col", "nhwc",
                                "upsample", "concat", "timestep", "scale_model", "euler", "advance")):
            name = "(torch) " + name[:40]
        d = per.setdefault(name, {"ids": set(), "time_ns": 0.0, "dram_read": 0.0, "dram_write": 0.0})
        d["ids"].add(r["ID"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["time_ns"] += v * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "msecond": 1e6, "ms": 1e6}.get(unit, 1)
        elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d["dram_read" if "read" in m else "dram_write"] += v * scale
    total = sum(d["time_ns"] for d in per.values())
    out = {"source": path, "total_time_ms": total / 1e6, "families": {}}
    print(f"{'kernel':34s} {'launches':>8s} {'time ms':>9s} {'share':>6s} {'dram rd MB':>11s} {'dram wr MB':>11s}")
    for name, d in sorted(per.items(), key=lambda kv: -kv[1]["time_ns"]):
        n = len(d["ids"])
        print(f"{name:34s} {n:8d} {d['time_ns'] / 1e6:9.3f} {d['time_ns'] / total:6.1%} {d['dram_read'] / 1e6:11.1f} "
              f"{d['dram_write'] / 1e6:11.1f}")
        out["families"][name] = {"launches": n, "time_ms": d["time_ns"] / 1e6, "share": d["time_ns"] / total,
                                 "dram_read_bytes": d["dram_read"], "dram_write_bytes": d["dram_write"]}
    print(f"total {total / 1e6:.3f} ms over {sum(len(d['ids']) for d in per.values())} launches (serialised, cold clocks)")
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
