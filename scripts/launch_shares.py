"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: counts, totals, shares."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]; ik = h.index("Kernel Name"); iv = h.index("Metric Value"); iu = h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= iv: continue
    name = r[ik].split("(")[0][:60]
    try: v = float(r[iv].replace(",", ""))
    except ValueError: continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    agg[name][0] += 1; agg[name][1] += v * scale
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':62s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} {v[0]:8d} {v[1]/1e3:10.3f} {v[1]/tot*100:6.1f}% {v[1]/v[0]:9.1f}")
print(f"{'total':62s} {sum(v[0] for v in agg.values()):8d} {tot/1e3:10.3f}")
