#!/bin/bash
# Refresh profiles/ from the latest gpurun_out/ capture (run on the CPU box).
set -e
cd "$(dirname "$0")/.."
TAG=${1:-r02}
export TAG
FR=${2:-4096}
export FR
python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep gpurun_out/launches.csv profiles/${TAG}_ncu_c3_${FR}frames.md \
  "Round ${TAG#r0} — ncu, bench.py --frames ${FR} --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra (C3, $((FR*131072)) points/launch), 1x B200"
cp gpurun_out/launches.csv profiles/${TAG}_launches_c3_${FR}frames.csv
python - <<'PY'
import csv, io, json, subprocess
out = subprocess.run(["ncu", "-i", "gpurun_out/prof.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u = rows[0], rows[1]
r = next(r for r in rows[2:] if "k_points_pair" in r[h.index("Kernel Name")] or "k_points<" in r[h.index("Kernel Name")])
def val(k):
    v = float(r[h.index(k)].replace(",", "")); unit = u[h.index(k)]
    return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[unit]
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
import os
fr = int(os.environ.get('FR', '4096'))
tag = os.environ.get('TAG', 'r02')
pts = fr * 131072
json.dump({"kernel": r[h.index("Kernel Name")], "points_per_launch": pts, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_point": (rd + wr) / pts,
           "source": f"profiles/{tag}_ncu_c3_{fr}frames.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one k_points_pair launch over {pts} points)"},
          open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print(open(f"profiles/{tag}_traffic.json").read())
PY
