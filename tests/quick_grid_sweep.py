"""Tuning script (not a test): bench.py phase bandwidths for different streaming-grid caps (LBFGS_B200_CTAS_PER_SM)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for cap in sys.argv[1:] or ["8", "4", "6", "2"]:
    env = dict(os.environ, LBFGS_B200_CTAS_PER_SM=cap)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline"],
                       env=env, capture_output=True, text=True)
    try:
        line = json.loads(r.stdout.strip().splitlines()[-1])
        print(cap, "value %.1f" % line["value"], "e2e %.1f" % line["e2e"]["value"],
              {k: round(v) for k, v in line["phase_gb_per_s"].items()}, {k: round(v, 3) for k, v in line["phase_ms_per_step"].items()},
              "hv_full %.0f" % line["roofline"]["full_history"]["gb_per_s"], flush=True)
    except Exception as e:  # noqa: BLE001
        print(cap, "failed", e, r.stdout[-500:], r.stderr[-1500:], flush=True)
