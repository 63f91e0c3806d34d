#!/bin/bash
# Refresh profiles/ from the latest gpurun_out/ capture (run on the CPU box).
set -e
cd "$(dirname "$0")/.."
TAG=${1:-r01}
python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep gpurun_out/launches.csv profiles/${TAG}_ncu_c3_512frames.md \
  "Round 1 — ncu, bench.py --frames 512 --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra (C3 geometry, 67.1M points/launch), 1x B200"
cp gpurun_out/launches.csv profiles/${TAG}_launches_c3_512frames.csv
python - <<'PY'
import csv, io, json, subprocess
out = subprocess.run(["ncu", "-i", "gpurun_out/prof.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u = rows[0], rows[1]
r = next(r for r in rows[2:] if "k_points" in r[h.index("Kernel Name")])
def val(k):
    v = float(r[h.index(k)].replace(",", "")); unit = u[h.index(k)]
    return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[unit]
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
pts = 512 * 131072
json.dump({"kernel": r[h.index("Kernel Name")], "points_per_launch": pts, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_point": (rd + wr) / pts,
           "source": "profiles/r01_ncu_c3_512frames.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, 67.1M-point launch, scaled per point)"},
          open("profiles/r01_traffic.json", "w"), indent=1)
print(open("profiles/r01_traffic.json").read())
PY
