#!/bin/bash
# profiles/capture.sh -- regenerate the ncu evidence of profiles/ on a GPU box (one GPU; ncu replays every kernel ~40 times):
#     gpurun --timeout 1500 -- 'bash profiles/capture.sh r02'
# writes gpurun_out/<round>_*.ncu-rep and the launch list; afterwards, in the build container:
#     bash profiles/capture.sh r02 export        # .ncu-rep -> profiles/<round>_*.{raw,details}.csv + <round>_*_hotspots.txt
# The command profiled is `python bench.py --config cN --steps 1 --warmup 3 --profile-mode` (the 4th launch of k_persist = one
# whole minimize(): C2 22 iterations / 71 rounds).  Numbers printed by runs under ncu are never bench values.
set -u
ROUND=${1:-rXX}
MODE=${2:-capture}
cd "$(dirname "$0")/.."
OUT=gpurun_out
if [ "$MODE" = "capture" ]; then
    mkdir -p $OUT
    CMD="python bench.py --config c2 --steps 1 --warmup 3 --profile-mode --no-cpu-baseline"
    ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${ROUND}_launches_bench_c2.csv $CMD > $OUT/${ROUND}_launches.log 2>&1
    for c in c2 c3; do
        ncu --set full --clock-control none --import-source on -k regex:k_persist -s 3 -c 1 -o $OUT/${ROUND}_k_persist_${c}_full \
            python bench.py --config $c --steps 1 --warmup 3 --profile-mode --no-cpu-baseline > $OUT/${ROUND}_k_persist_${c}.log 2>&1
        tail -2 $OUT/${ROUND}_k_persist_${c}.log | cut -c1-200
    done
    ls -la $OUT/${ROUND}_*
else
    for rep in $OUT/${ROUND}_*_full.ncu-rep; do
        base=profiles/$(basename "${rep%.ncu-rep}")
        ncu -i "$rep" --page raw --csv > "$base.raw.csv" 2>/dev/null
        ncu -i "$rep" --page details --csv > "$base.details.csv" 2>/dev/null
        ncu -i "$rep" --page source --csv --print-source cuda,sass > /tmp/_src.csv 2>/dev/null
        python profiles/hotspots.py /tmp/_src.csv "$base.raw.csv" > "${base}_hotspots.txt"
        echo "exported $base.{raw,details}.csv ${base}_hotspots.txt"
    done
    [ -f $OUT/${ROUND}_launches_bench_c2.csv ] && cp $OUT/${ROUND}_launches_bench_c2.csv profiles/
    echo "now update profiles/README.md and profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of the k_persist launch)"
fi
