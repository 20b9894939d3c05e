"""DRAM traffic per launch of one kernel from an `ncu --set full` capture, as the JSON bench.py reads.

    ncu -i capture.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_traffic.py raw.csv "rowgemm[dy/spatial_bwd]" "<how the capture was made>" > profiles/rN_traffic_X.json

dram_bytes_per_launch = mean over the captured launches of dram__bytes_read.sum + dram__bytes_write.sum."""
import csv, json, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}


def val(d, key):
    v, u = float(d[idx[key]].replace(",", "")), units[idx[key]]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3,
             "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
    return v * scale[u]


per = []
for d in data:
    rd, wr = val(d, "dram__bytes_read.sum"), val(d, "dram__bytes_write.sum")
    per.append([d[idx["Kernel Name"]][:80], round(rd / 1e6, 1), round(wr / 1e6, 1), round(val(d, "gpu__time_duration.sum"), 1)])
total = sum(p[1] + p[2] for p in per) * 1e6
print(json.dumps({"kernel": sys.argv[2], "source": sys.argv[3], "dram_bytes_per_launch": total / len(per),
                  "launches": len(per), "per_launch_MB_read_write_us": per}, indent=1))
