"""Dev tool: condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-(kernel, grid) totals.
    python tools/launch_summary.py gpurun_out/x_launches.csv "<command that was profiled>" > profiles/x.txt
"""
import collections
import csv
import sys


def main(path, cmd=""):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0][:80]
        a = agg.setdefault((name, r["Grid Size"], r["Block Size"]), [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    print(f"ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches): {cmd}")
    print(f"launches captured: {len(rows)}  total {tot / 1e6:.3f} ms")
    for (n, g, b), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e6:10.3f} ms {100 * t / tot:6.2f}% {c:4d}x {t / c / 1e3:10.1f} us  {n}  grid={g} block={b}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
