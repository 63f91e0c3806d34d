#!/bin/bash
# ncu launch list (per-kernel durations) of the default bench at full size
mkdir -p gpurun_out
BENCH_ARGS="--frames ${PFRAMES:-4096} --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra"
python bench.py $BENCH_ARGS > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function --kernel-name regex:'^k_' -c 400 \
    --csv --log-file gpurun_out/launches.csv python bench.py $BENCH_ARGS > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows[-40:]:
    print(r[4][:60], r[-1])
PY
