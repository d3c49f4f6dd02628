"""Per-instruction view of an `ncu --page source --csv --print-source sass` dump: for the address range that holds most
samples, print each SASS instruction with its sample count, executions and dominant stall reasons.  Development aid.
  python scripts/ncu_sass_hot.py dump.csv [min_exec_fraction]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        samp = int(r[ix["# Samples"]])
        ex = int(r[ix["Instructions Executed"]])
    except ValueError:
        continue
    st = {h: int(r[ix[h]] or 0) for h in stall_cols}
    data.append((r[ix["Address"]], r[ix["Source"]].strip(), samp, ex, st))
tot = sum(d[2] for d in data) or 1
mx = max(d[3] for d in data) or 1
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
hot = [d for d in data if d[3] >= frac * mx]
print(f"total samples {tot}; instructions with executions >= {frac:.2f} * max ({mx}): {len(hot)}, holding {sum(d[2] for d in hot) / tot * 100:.1f}% of samples")
for a, s, samp, ex, st in hot:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    tops = " ".join(f"{k[6:]}={v}" for k, v in top if v)
    print(f"{samp:7d} {samp / tot * 100:5.2f}%  x{ex / mx:4.2f}  {s[:70]:70s} {tops}")
