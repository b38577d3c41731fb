#!/bin/bash
# profiles/env_sweep.sh <round> <config> "<ENV=val ...>" ... -- one bench line (no CPU leg) per environment setting
# switches of the shipped library: LBFGS_B200_STAGES (2|3|4 stages of the staging ring), LBFGS_B200_TUNE (4 = evict-first trial stores),
# LBFGS_B200_VIRTUAL_FIRST_TRIAL, LBFGS_B200_CTAS_PER_SM (host-driven loop).  LBFGS_B200_SPECULATE existed at commit 2f7dc75 only
# (DESIGN.md section 10); the r02c/r02d/r02e lines under r02_experiments/ were taken with that build.
ROUND=$1; CFG=$2; shift 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
i=0
for envs in "$@"; do
    i=$((i+1))
    tag=$(echo "$envs" | tr ' =' '__' | tr -cd 'A-Za-z0-9_')
    steps=10; [ "$CFG" = "c5" ] && steps=2
    env $envs timeout 300 python bench.py --config $CFG --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/${ROUND}_${CFG}_$tag.json 2> gpurun_out/${ROUND}_${CFG}_$tag.err
    python - <<P
import json
try:
    d = json.loads(open("gpurun_out/${ROUND}_${CFG}_$tag.json").read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print("$CFG [$envs] value %.1f it/s  ms %.3f  e2e %.1f  frac %s  sync %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("frac"), r.get("sync_ms_per_solve")),
          d.get("setup", {}).get("step"), {k: (round(v["ms_per_solve"], 3), v["rounds_per_solve"], round(v["gb_per_s"] or 0)) for k, v in (r.get("passes") or {}).items()})
except Exception as e:
    print("$CFG [$envs] parse failed", e); print(open("gpurun_out/${ROUND}_${CFG}_$tag.err").read()[-1500:])
P
done
