#!/bin/bash
# Full-size bench + ncu launch list (+ one --set full capture of the top kernels). Logs -> gpurun_out/.
mkdir -p gpurun_out
BENCH_ARGS="--frames ${PFRAMES:-4096} --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra"
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench_full exit $?"
tail -1 gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench_ref exit $?"
tail -1 gpurun_out/bench_ref.log | cut -c1-300
python bench.py $BENCH_ARGS > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function --kernel-name regex:'^k_' -c 400 \
    --csv --log-file gpurun_out/launches.csv python bench.py $BENCH_ARGS > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
if [ "${FULL:-1}" = "1" ]; then
python bench.py $BENCH_ARGS > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base function --kernel-name regex:'^k_(points_pair|points_deferred|sweep_compact|sweep_walk|miss_fold|finalize)' -s ${SKIP:-7} -c 7 \
    -o gpurun_out/prof -f python bench.py $BENCH_ARGS > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null
ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source sass --kernel-name-base function --kernel-name regex:'^k_points_pair' > gpurun_out/src_pair.csv 2>/dev/null
ncu -i gpurun_out/prof.ncu-rep --page raw --csv --kernel-name-base function --kernel-name regex:'^k_points_pair' > gpurun_out/prof_pair_raw.csv 2>/dev/null
fi
ls -la gpurun_out/
