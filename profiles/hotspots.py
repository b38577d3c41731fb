"""Summarise an ncu capture: duration / DRAM bytes from the raw page, and the source lines that collect the most warp samples
(with their stall reasons) from the `--page source --print-source cuda,sass` export.  Usage: hotspots.py <source.csv> <raw.csv>"""
import csv
import sys


def main(src_csv, raw_csv):
    rows = list(csv.reader(open(raw_csv)))
    if len(rows) >= 3:
        hdr, units, vals = rows[0], rows[1], rows[2]
        want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size")
        for k, u, v in zip(hdr, units, vals):
            if k in want:
                print("%-60s %s %s" % (k, v, u))
    print()
    rows = list(csv.reader(open(src_csv)))
    cur_file, ix, data = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            ix = {}
            for i, k in enumerate(r):
                ix.setdefault(k, i)
            continue
        if r[0] != "" and ix:
            try:
                s = int(r[ix["# Samples"]])
            except (ValueError, KeyError):
                continue

            def g(k):
                try:
                    return int(r[ix[k]])
                except (ValueError, KeyError):
                    return 0
            data.append((s, cur_file, r[0], r[1].strip()[:90], g("stall_barrier"), g("stall_long_sb"), g("stall_short_sb"), g("stall_wait"), g("stall_lg")))
    tot = sum(d[0] for d in data) or 1
    print("warp samples by source line (all warps of all CTAs; idle warps waiting at a barrier count too): total %d" % tot)
    print("%7s %6s  %-22s %8s %8s %8s %8s %8s  source" % ("samples", "share", "file:line", "barrier", "long_sb", "short_sb", "wait", "lg"))
    for d in sorted(data, reverse=True)[:30]:
        print("%7d %5.1f%%  %-22s %8d %8d %8d %8d %8d  %s" % (d[0], 100.0 * d[0] / tot, "%s:%s" % (d[1][:16], d[2]), d[4], d[5], d[6], d[7], d[8], d[3]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
