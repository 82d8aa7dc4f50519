timeout 120 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:resize_ -c 60 --csv --log-file gpurun_out/r02j_resize_launches.csv python tools/bench_resize.py 3 4 6 > /dev/null 2>&1; echo "list rc=$?"
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02j_resize_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID'); gi=h.index('Grid Size'); bi=h.index('Block Size')
d={}
for r in rows[1:]: d.setdefault((int(r[ii]),r[ki][:44],r[gi],r[bi]),{})[r[mi][:24]]=r[vi]
for k in sorted(d):
    if k[0] % 16 in (4,5): print(k, d[k])
PY
