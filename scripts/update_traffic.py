"""profiles/traffic.json <- dram__bytes_read.sum + dram__bytes_write.sum of one captured launch.
Usage: python scripts/update_traffic.py <ncu-rep> <key> <source text>"""
import csv, io, json, os, subprocess, sys
rep, key, source = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name)
    v, u = float(vals[i]), units[i]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
tr = json.load(open(path))
tr[key] = {"dram_bytes_per_launch": int(rd + wr), "source": f"{source}: dram__bytes_read.sum {rd / 1e9:.3f} GB + dram__bytes_write.sum {wr / 1e9:.3f} GB"}
json.dump(tr, open(path, "w"), indent=2)
print(key, tr[key])
