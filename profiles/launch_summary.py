"""Aggregate an ncu --metrics gpu__time_duration.sum launch list per kernel: count, total, mean, share."""
import csv, sys, re, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); ig = hdr.index("Grid Size"); ib = hdr.index("Block Size")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("<unnamed>::", "").replace("void ", "")
    key = (name, r[ig], r[ib])
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += float(r[iv].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':50s} {'grid':>14s} {'block':>12s} {'n':>5s} {'total_us':>10s} {'mean_us':>9s} {'share':>6s}")
for (name, g, b), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:50s} {g:>14s} {b:>12s} {n:5d} {t/1e3:10.1f} {t/1e3/n:9.2f} {100*t/tot:5.1f}%")
print(f"total {tot/1e3:.1f} us over {sum(a[0] for a in agg.values())} launches")
