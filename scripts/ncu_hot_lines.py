"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line (development aid)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr=None; cur_file=None; agg={}
def num(x):
    try: return int(float(x))
    except: return 0
for r in rows:
    if r and r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r and r[0]=='Line No': hdr=r; ia=hdr.index("Instructions Executed"); isamp=hdr.index("# Samples"); continue
    if hdr and r and r[0] not in ('','Function Name'):
        try: ln=int(r[0])
        except: continue
        key=(cur_file,ln,r[1].strip()[:100])
        a,s = agg.get(key,(0,0))
        agg[key]=(a+num(r[ia]), s+num(r[isamp]))
tot=sum(v[0] for v in agg.values()) or 1; ts=sum(v[1] for v in agg.values()) or 1
print("total inst", tot, "samples", ts)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:topn]:
    print(f"{v[0]/tot*100:5.1f}% inst {v[1]/ts*100:5.1f}% samp  {k[0]}:{k[1]}  {k[2]}")
