#!/bin/bash
# profiles/run_multi_gpu_bench.sh <round> -- on an 8-GPU box: gpurun --gpus 8 --timeout 900 -- 'bash profiles/run_multi_gpu_bench.sh r02'
# one bench line per sharded configuration (config 2 n-sharded; config 5 problem-parallel and n-sharded)
ROUND=${1:-rXX}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name nproc env args...
  name=$1; np=$2; envs=$3; shift 3
  env $envs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "== $name rc=$?"; tail -c 300 gpurun_out/$name.err | tail -2
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/$name.json").read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print(d["n_gpus"], "value %.1f  ms/step %.3f  e2e %.1f  sync %s  wait_last %s  exchange %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"],
          r.get("sync_ms_per_solve"), r.get("sync_wait_last_cta_ms_per_solve"), r.get("sync_cross_rank_exchange_ms_per_solve")), json.dumps(d.get("setup"))[:300])
except Exception as e:
    print("parse failed", e)
P
}
N=${2:-8}
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${ROUND}_gpu_multi_suite.log 2>&1; echo "2-GPU parity suite rc=$?"; tail -2 gpurun_out/${ROUND}_gpu_multi_suite.log
run ${ROUND}_bench_c2_n$N $N "A=0" --steps 10 --warmup 3
run ${ROUND}_bench_c5_n${N}_problems $N "A=0" --config c5 --sharding problems --steps 2 --warmup 1
run ${ROUND}_bench_c5_n${N}_nsharded $N "A=0" --config c5 --sharding n --steps 2 --warmup 1
