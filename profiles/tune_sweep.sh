#!/bin/bash
# profiles/tune_sweep.sh <round> <tune values...> -- bench c2 (no CPU leg) under LBFGS_B200_TUNE=<v>: it/s and the per-pass GB/s
ROUND=$1; shift
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "$@"; do
    LBFGS_B200_TUNE=$v timeout 300 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${ROUND}_tune_$v.json 2> gpurun_out/${ROUND}_tune_$v.err
    python - <<P
import json
try:
    d = json.loads(open("gpurun_out/${ROUND}_tune_$v.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("tune=$v value %.1f it/s  ms %.3f  frac %.3f  sync %.3f " % (d["value"], d["ms_per_step"], r["frac"], r["sync_ms_per_solve"]),
          {k: (round(v["ms_per_solve"], 3), round(v["gb_per_s"])) for k, v in r["passes"].items()})
except Exception as e:
    print("tune=$v parse failed", e)
P
done
