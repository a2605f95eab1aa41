"""Aggregate an ncu launch-list CSV (gpu__time_duration.sum) by kernel name."""
import collections, csv, sys
for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try: v = float(r[vi].replace(",", ""))
        except ValueError: continue
        if r[ui] == "ns": v /= 1000
        a = agg.setdefault(r[ki][:56], [0, 0.0]); a[0] += 1; a[1] += v
    print(path)
    for k, (n, t) in agg.items(): print(f"  {k:56s} n={n:3d} avg={t / n:8.1f} us")
