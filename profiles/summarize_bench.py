"""Markdown table of the committed bench lines (profiles/<round>_bench_*.json, one JSON line each, written by bench.py).
Usage: python profiles/summarize_bench.py r02"""
import glob
import json
import os
import sys


def main(rnd):
    here = os.path.dirname(os.path.abspath(__file__))
    rows = []
    for path in sorted(glob.glob(os.path.join(here, rnd + "_bench_*.json"))):
        try:
            d = json.loads(open(path).read().strip().splitlines()[-1])
        except Exception:
            continue
        r = d.get("roofline", {})
        cb = d.get("cpu_baseline", {})
        rows.append((os.path.basename(path), d.get("n_gpus"), d.get("value"), d.get("ms_per_step"), d.get("e2e", {}).get("value"),
                     r.get("achieved"), r.get("frac"), r.get("sync_ms_per_solve"), cb.get("value"), d.get("launches_per_solve")))
    print("| file | GPUs | it/s (HBM-resident) | ms / solve | it/s end to end | roofline GB/s | frac of copy peak | sync ms / solve | CPU 1 core it/s | launches / solve |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    f = lambda v, p="%.1f": "-" if v is None else p % v
    for r in rows:
        print("| %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (r[0], r[1], f(r[2]), f(r[3], "%.3f"), f(r[4]), f(r[5], "%.0f"), f(r[6], "%.3f"),
                                                                     f(r[7], "%.3f"), f(r[8], "%.2f"), f(r[9], "%.0f")))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
