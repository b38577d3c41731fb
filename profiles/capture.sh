#!/bin/bash
# profiles/capture.sh -- regenerate the ncu evidence of profiles/ on a GPU box (one GPU; ncu replays every kernel ~40 times):
#     gpurun --timeout 1500 -- 'bash profiles/capture.sh r02'
# writes gpurun_out/<round>_*.ncu-rep and the launch list; afterwards, in the build container:
#     bash profiles/capture.sh r02 export        # .ncu-rep -> profiles/<round>_*.{raw,details}.csv
# The command profiled is always `python bench.py --steps 1 --warmup 3 --profile-mode` (config C2: 22 iterations per solve, the
# history is full (c = 10) from the 11th iteration on).  Numbers printed by runs under ncu are never bench values.
set -u
ROUND=${1:-rXX}
MODE=${2:-capture}
cd "$(dirname "$0")/.."
OUT=gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --profile-mode"
# kernel regex : launches to skip so that the captured one runs with a full history in the 4th solve
#   fused calls per solve: 21 (k_pair_dots, k_gram_combine); trials per solve: 50; 3 warm-up solves come first
declare -A SKIP=( [k_pair_dots]=75 [k_gram_combine]=75 [k_trial]=165 [k_gram_dots]=0 [k_update]=0 )
if [ "$MODE" = "capture" ]; then
    mkdir -p $OUT
    ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 80 --csv --log-file $OUT/${ROUND}_launches_bench_c2.csv $CMD > $OUT/${ROUND}_launches.log 2>&1
    for k in k_pair_dots k_gram_combine k_trial; do
        ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP[$k]} -c 1 -o $OUT/${ROUND}_${k}_c10_full $CMD > $OUT/${ROUND}_${k}.log 2>&1
        tail -2 $OUT/${ROUND}_${k}.log | cut -c1-200
    done
    ls -la $OUT/${ROUND}_*
else
    for rep in $OUT/${ROUND}_*_full.ncu-rep; do
        base=profiles/$(basename "${rep%.ncu-rep}")
        ncu -i "$rep" --page raw --csv > "$base.raw.csv" 2>/dev/null
        ncu -i "$rep" --page details --csv > "$base.details.csv" 2>/dev/null
        echo "exported $base.{raw,details}.csv"
    done
    [ -f $OUT/${ROUND}_launches_bench_c2.csv ] && cp $OUT/${ROUND}_launches_bench_c2.csv profiles/
    echo "now update profiles/README.md and profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of k_pair_dots + k_gram_combine)"
fi
