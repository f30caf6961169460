python tools/bench_c3.py 1.0 2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c3_r2b.csv python tools/bench_c3.py 1.0 2 > /dev/null 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/c3_r2b.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[hdr]; ki=H.index("Kernel Name"); vi=H.index("Metric Value")
names=[(r[ki],float(r[vi].replace(',',''))) for r in rows[hdr+1:] if len(r)>vi]
idx=[i for i,(n,v) in enumerate(names) if 'march' in n]
a,b=idx[-2],idx[-1]
agg=collections.OrderedDict(); tot=0
for n,v in names[a:b]:
    k=n[:60]; agg.setdefault(k,[0,0]); agg[k][0]+=v; agg[k][1]+=1; tot+=v
for k,(v,c) in agg.items():
    if v>50e3: print(f"{v/1e3:9.1f} us x{c:3d}  {k}")
print("sum", tot/1e6, "ms")
PY
