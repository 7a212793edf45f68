"""Developer tool: aggregate an `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list per kernel
(launches, total and mean duration, share of the listed time) into the csv committed under profiles/.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-baselines
    python tools/ncu_launches.py gpurun_out/launches.csv profiles/r2_ncu_launches_bench.csv "<note>"
"""
import csv
import sys


def main():
    src, dst, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    rows = list(csv.reader(open(src, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Value" in r)
    hdr = rows[hi]
    ni, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    agg = {}
    for r in rows[hi + 1:]:
        if len(r) != len(hdr) or r[mi] != "gpu__time_duration.sum":
            continue
        us = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        c, t = agg.get(r[ni], (0, 0.0))
        agg[r[ni]] = (c + 1, t + us)
    tot = sum(t for _, t in agg.values())
    with open(dst, "w") as f:
        f.write(f'"# {note}"\n')
        f.write("kernel,launches,total_us,share_pct,us_per_launch\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"' + k[:100].replace('"', "'") + f'",{c},{t:.1f},{100 * t / tot:.2f},{t / c:.2f}\n')
    print(open(dst).read())


if __name__ == "__main__":
    main()
