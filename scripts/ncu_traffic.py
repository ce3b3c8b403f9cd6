"""Extract per-launch DRAM traffic of a kernel from an ncu report into profiles/traffic.json.
usage: python scripts/ncu_traffic.py <report.ncu-rep> <key> <batch>"""
import csv, json, os, subprocess, sys
rep, key, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name); v = float(vals[i]); u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d[key] = {"batch": batch, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
          "kernel": vals[hdr.index("Kernel Name")], "duration_us_under_ncu": get("gpu__time_duration.sum") / (1e3 if units[hdr.index("gpu__time_duration.sum")] == "ns" else 1),
          "source": os.path.basename(rep)}
json.dump(d, open(path, "w"), indent=1)
print(d[key])
