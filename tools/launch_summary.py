"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, mean, share."""
import collections, csv, sys

for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r is hdr:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("agcf::", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("## %s -- %d launches, %.1f us of kernel time" % (f, sum(v[0] for v in agg.values()), tot / 1e3))
    print("| kernel | launches | total us | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.2f | %.1f %% |" % (k[:90], v[0], v[1] / 1e3, v[1] / v[0] / 1e3, 100 * v[1] / tot))
    print()
