#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:vp8|gif|resize_area|compact|extract_alpha|vp8l" -c 300 --csv --log-file $O/r02_launches_c4.csv \
  python bench.py --config 4 --batch 32 --steps 1 --warmup 1 --no-cpu-baseline > $O/prof_c4_list.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:vp8_encode_batch_kernel" --launch-skip 1 -c 1 -o $O/r02_vp8_encode -f \
  python bench.py --config 4 --batch 32 --steps 1 --warmup 1 --no-cpu-baseline > $O/prof_c4_full.log 2>&1; echo "full rc=$?"
python - <<'P'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02_launches_c4.csv')))
i=[k for k,r in enumerate(rows) if 'Kernel Name' in r][0]; h=rows[i]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[i+1:]:
    if len(r)<=mv: continue
    a=agg.setdefault(r[kn].split('(')[0],[0,0.0]); a[0]+=1; a[1]+=float(r[mv].replace(',',''))
for k,v in agg.items(): print(k,v[0],'launches',round(v[1]/1e6,2),'ms')
P
