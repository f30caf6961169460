python tools/bench_train.py --steps 5 --warmup 2 2>&1 | tail -2 | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_r2a.csv python tools/bench_train.py --steps 2 --warmup 0 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/train_r2a.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[hdr]; ki=H.index("Kernel Name"); vi=H.index("Metric Value")
names=[(r[ki],float(r[vi].replace(',',''))) for r in rows[hdr+1:] if len(r)>vi]
idx=[i for i,(n,v) in enumerate(names) if 'march' in n]
a,b=idx[-2],idx[-1]
tot=0
import collections
agg=collections.OrderedDict()
for n,v in names[a:b]:
    k=n[:70]; agg.setdefault(k,[0,0]); agg[k][0]+=v; agg[k][1]+=1; tot+=v
for k,(v,c) in agg.items(): print(f"{v/1e3:9.1f} us x{c:3d}  {k}")
print("sum", tot/1e6, "ms", len(names[a:b]), "launches")
PY
ncu --set full --import-source on --clock-control none -k regex:render_composite -c 2 -o gpurun_out/tail_r2a python bench.py --steps 1 --warmup 3 --no-strong --no-reference-gpu --no-cpu-baseline --no-semantic-variant --no-train-step > /dev/null 2>&1
ls -la gpurun_out/tail_r2a.ncu-rep
