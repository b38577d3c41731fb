#!/bin/bash
# profiles/run_two_gpu.sh <round> -- on a 2-GPU box: gpurun --gpus 2 --timeout 900 -- 'bash profiles/run_two_gpu.sh r02'
# the 2-GPU parity tests (in-kernel NVLink exchange and NCCL), then one bench line per sharded configuration.
set -u
ROUND=${1:-rXX}
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/${ROUND}_gpu_multi_suite.log 2>&1; echo "multi suite rc=$?"; tail -3 $OUT/${ROUND}_gpu_multi_suite.log
run() { # name nproc args...
  name=$1; np=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $np "$@" > $OUT/$name.json 2> $OUT/$name.err
  echo "== $name rc=$?"; tail -c 300 $OUT/$name.err | tail -2
  python - <<P
import json
try:
    d=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print(d["n_gpus"], "value %.1f  ms/step %.3f  e2e %.1f  sync %s  wait_last %s  exchange %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"],
          r.get("sync_ms_per_solve"), r.get("sync_wait_last_cta_ms_per_solve"), r.get("sync_cross_rank_exchange_ms_per_solve")), json.dumps(d.get("setup"))[:300])
except Exception as e:
    print("parse failed", e)
P
}
run ${ROUND}_bench_c2_n2 2 --steps 10 --warmup 3
run ${ROUND}_bench_c3_n2 2 --config c3 --steps 10 --warmup 3
run ${ROUND}_bench_c5_n2_problems 2 --config c5 --sharding problems --steps 2 --warmup 1
run ${ROUND}_bench_c5_n2_nsharded 2 --config c5 --sharding n --steps 2 --warmup 1
