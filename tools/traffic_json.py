"""DRAM traffic of one captured launch against its algorithmic bytes -> the small JSON files bench.py reads for `roofline.traffic`.
usage: python tools/traffic_json.py x.ncu-rep "<kernel label>" <algorithmic bytes> [--index K] > profiles/r2_traffic_*.json"""
import csv, io, json, subprocess, sys
rep, label, algo = sys.argv[1], sys.argv[2], float(sys.argv[3])
K = int(sys.argv[sys.argv.index("--index") + 1]) if "--index" in sys.argv else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + K]
def get(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(units[i], 1)
rd, wr, us = get("dram__bytes_read.sum"), get("dram__bytes_write.sum"), get("gpu__time_duration.sum")
print(json.dumps({"kernel": label, "capture": rep.split("/")[-1], "kernel_name": vals[hdr.index("Kernel Name")],
                  "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr, "algorithmic_bytes": algo,
                  "dram_bytes_per_algorithmic_byte": (rd + wr) / algo, "kernel_us": us,
                  "kernel_GBps_algorithmic": algo / us / 1e3}, indent=1))
