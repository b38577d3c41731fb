#!/bin/bash
# profiles/run_single_gpu.sh <round> [quick] -- one GPU, on the box:  gpurun --timeout 1500 -- 'bash profiles/run_single_gpu.sh r02'
# smoke, the -m gpu suite, one bench line per BASELINE config (both solver loops for c2), the pinned reference arm, then the ncu
# evidence (profiles/capture.sh).  Everything lands in gpurun_out/; copy what is to be judged into profiles/.
set -u
ROUND=${1:-rXX}
QUICK=${2:-}
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader > $OUT/${ROUND}_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${ROUND}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${ROUND}_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${ROUND}_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -2 $OUT/${ROUND}_gpu_suite.log
bench() { # name args...
    name=$1; shift
    timeout 400 python bench.py "$@" > $OUT/${ROUND}_bench_$name.json 2> $OUT/${ROUND}_bench_$name.err
    echo "== $name rc=$?"; tail -c 400 $OUT/${ROUND}_bench_$name.err | tail -2
    python - <<P
import json
try:
    d = json.loads(open("$OUT/${ROUND}_bench_$name.json").read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print("value %.1f it/s  %.3f ms/step  e2e %.1f  roofline %.0f GB/s frac %.3f  other %s  cpu %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("achieved", 0), r.get("frac", 0),
        json.dumps(d.get("other_loop", {}).get("value")), json.dumps(d.get("cpu_baseline", {}).get("value"))))
except Exception as e:
    print("parse failed", e)
P
}
bench c2_n1 --config c2 --steps 20 --warmup 3
bench c3_n1 --config c3 --steps 20 --warmup 3
bench c5_n1 --config c5 --steps 2 --warmup 1
bench c4_n1 --config c4 --steps 10 --warmup 3
if [ -z "$QUICK" ]; then
    bench c2_n1_hostloop --config c2 --solver-loop host --steps 20 --warmup 3 --no-cpu-baseline
    bench c2_reference --impl reference --config c2 --steps 2 --warmup 1
    timeout 900 bash profiles/capture.sh $ROUND
fi
ls -la $OUT | tail -30
