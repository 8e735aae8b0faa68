"""Summarise an ncu launch list into profiles/traffic_<tag>.json (per kernel: launches, time, DRAM bytes per launch).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 \
        --csv --log-file gpurun_out/launches_<tag>.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras
    python tools/ncu_traffic.py gpurun_out/launches_<tag>.csv profiles/traffic_<tag>.json "<command line>"

bench.py reads `kernels.conv_tc_kernel.dram_*_bytes_per_launch` for roofline.traffic.  Per-launch times under ncu are
cold-cache and serialised: use the kernel's SHARE of the listed time, not the absolute."""
import csv
import json
import re
import sys
from collections import defaultdict

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6,
        "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def main(src, dst, command):
    rows = []
    with open(src) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    per = defaultdict(lambda: {"launches": set(), "time_ms": 0.0, "rd": 0.0, "wr": 0.0})
    compact = {}    # id -> [kernel, grid, block, ns, read, write]   (written next to the JSON as a compact launch list)
    for r in rows:
        if "Kernel Name" in r:      # raw `ncu --csv` output: one row per (launch, metric)
            name = re.sub(r"<.*", "", r["Kernel Name"].split("(")[0]).split("::")[-1].strip()
            val = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1)
            m = r["Metric Name"]
            c = compact.setdefault(r["ID"], [name, r.get("Grid Size", ""), r.get("Block Size", ""), 0, 0, 0])
            k = per[name]
            k["launches"].add(r["ID"])
            if m.startswith("gpu__time_duration"):
                k["time_ms"] += val
                c[3] = int(val * 1e6)
            elif m.startswith("dram__bytes_read"):
                k["rd"] += val
                c[4] = int(val)
            elif m.startswith("dram__bytes_write"):
                k["wr"] += val
                c[5] = int(val)
        else:                       # compact list: ID,Kernel,Grid,Block,time [ns],read [B],write [B]
            keys = list(r.keys())
            name = re.sub(r"<.*", "", r["Kernel"]).strip()
            k = per[name]
            k["launches"].add(r["ID"])
            k["time_ms"] += float(r[keys[4]]) * 1e-6
            k["rd"] += float(r[keys[5]])
            k["wr"] += float(r[keys[6]])
    if compact:
        with open(dst.replace(".json", "_launches.csv").replace("traffic_", "launches_"), "w") as f:
            f.write("ID,Kernel,Grid,Block,gpu__time_duration.sum [ns],dram__bytes_read.sum [B],dram__bytes_write.sum [B]\n")
            for i, c in sorted(compact.items(), key=lambda kv: int(kv[0])):
                f.write(f'{i},{c[0]},"{c[1]}","{c[2]}",{c[3]},{c[4]},{c[5]}\n')
    total = sum(k["time_ms"] for k in per.values()) or 1.0
    out = {"source": command, "kernels": {}}
    for name, k in sorted(per.items(), key=lambda kv: -kv[1]["time_ms"]):
        n = len(k["launches"])
        out["kernels"][name] = {"launches": n, "time_ms": round(k["time_ms"], 4),
                                "share_of_listed_time": round(k["time_ms"] / total, 4),
                                "dram_read_bytes_per_launch": int(k["rd"] / n), "dram_write_bytes_per_launch": int(k["wr"] / n)}
    json.dump(out, open(dst, "w"), indent=1)
    for name, v in list(out["kernels"].items())[:12]:
        print(f"{name:32s} x{v['launches']:4d} {v['time_ms']:9.3f} ms {100 * v['share_of_listed_time']:5.1f} %  "
              f"rd {v['dram_read_bytes_per_launch'] / 1e6:9.1f} MB wr {v['dram_write_bytes_per_launch'] / 1e6:9.1f} MB per launch")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
