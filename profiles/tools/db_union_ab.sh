python profiles/tools/db_time.py 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:db_union_dense --csv --log-file gpurun_out/db_warm_x.csv python profiles/tools/db_time.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/db_warm_x.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; vi=H.index('Metric Value')
v=[float(r[vi].replace(',',''))/1000 for r in rows[hdr+2:] if len(r)>vi]
print('union warm us: frame1',sum(v[3:13])/10,'frame7',sum(v[16:26])/10)
PY
