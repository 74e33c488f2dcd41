python -m pytest tests -m gpu -q -x 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/chain_sparse.csv python tools/profile_chain.py 3 1 > /dev/null 2>&1
python - <<'P'
import csv, collections, re
f="gpurun_out/chain_sparse.csv"
rows=list(csv.reader(l for l in open(f) if l.startswith('"')))
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg=collections.OrderedDict(); n=len(rows)-1
for r in rows[1+2*n//3:]:
    m=re.search(r"(k_[a-z_0-9]+)", r[ki]); nm=m.group(1) if m else r[ki][:40]
    v=float(r[vi].replace(",","")); v = v/1e3 if r[ui].startswith("n") else v
    a=agg.setdefault(nm,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in agg.items(): print("   %-28s n=%d total %.1f us"%(k,c,t))
P
