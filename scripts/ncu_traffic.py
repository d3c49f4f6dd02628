"""Per-matrix DRAM / L2 traffic of a kernel from an `ncu --set full` capture -> profiles/kernel_traffic.json (read by
bench.py's roofline.traffic).  python scripts/ncu_traffic.py <key> <capture.ncu-rep> <matrices in the captured launch>
e.g.  python scripts/ncu_traffic.py stageB_c64_n64 gpurun_out/r02e_stageB_c64.ncu-rep 2960"""
import csv
import io
import json
import os
import subprocess
import sys

key, rep, nmat = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def num(name):
    u, v = d[name]
    x = float(v.replace(",", ""))
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(root, "profiles", "kernel_traffic.json")
db = json.load(open(path)) if os.path.exists(path) else {}
sha = subprocess.run(["git", "-C", root, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
db[key] = {
    "kernel": d["Kernel Name"][1],
    "capture": os.path.basename(rep),
    "git_sha_at_extraction": sha,
    "matrices_in_launch": nmat,
    "dram_bytes_per_matrix": (num("dram__bytes_read.sum") + num("dram__bytes_write.sum")) / nmat,
    "dram_read_bytes_per_matrix": num("dram__bytes_read.sum") / nmat,
    "dram_write_bytes_per_matrix": num("dram__bytes_write.sum") / nmat,
    "l2_bytes_per_matrix": float(d["lts__t_sectors.sum"][1].replace(",", "")) * 32.0 / nmat,
    "gpu_time_ms": float(d["gpu__time_duration.sum"][1]),
}
json.dump(db, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(db[key], indent=1))
