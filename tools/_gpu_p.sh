ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/edit_r2b.csv python tools/bench_edit.py > /dev/null 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/edit_r2b.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[hdr]; ki=H.index("Kernel Name"); vi=H.index("Metric Value")
names=[(r[ki],float(r[vi].replace(',',''))) for r in rows[hdr+1:] if len(r)>vi]
idx=[i for i,(n,v) in enumerate(names) if 'cell_min_kernel' in n]
a=idx[-1]
b=[i for i,(n,v) in enumerate(names) if 'march' in n and i>a][0]
tot=0
for n,v in names[a-22:b]:
    if v>20e3: print(f"{v/1e3:9.1f} us  {n[:80]}")
    tot+=v
print("sum", tot/1e6)
PY
