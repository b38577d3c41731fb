run() { # name nproc args...
  name=$1; np=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "== $name rc=$?"; tail -c 300 gpurun_out/$name.err | tail -2
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/$name.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("roofline",{}).get("sync_ms_per_solve"), json.dumps(d.get("setup"))[:300])
except Exception as e:
    print("parse failed", e)
P
}
run r02_bench_c2_n8 8 --steps 10 --warmup 3
run r02_bench_c2_n4 4 --steps 10 --warmup 3
run r02_bench_c5_n8_problems 8 --config c5 --sharding problems --steps 2 --warmup 1
run r02_bench_c5_n8_nsharded 8 --config c5 --sharding n --steps 2 --warmup 1
