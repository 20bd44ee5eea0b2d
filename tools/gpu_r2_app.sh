#!/bin/bash
# where the GPU time of the application-level scan (ft2.phases, 7 x 10^6 rays) goes: ncu launch list
O=gpurun_out/r02
mkdir -p $O
timeout 600 python tools/app_profile.py 1000003 > $O/app_plain.txt 2>&1; head -3 $O/app_plain.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/app_launches.csv python tools/app_profile.py 1000003 > $O/app_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r02/app_launches.csv')))
hdr = None
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get('Metric Name') == 'gpu__time_duration.sum':
            v = float(d['Metric Value'].replace(',', '')); u = d['Metric Unit']
            v *= {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'nsecond': 1e-6}.get(u, 1e-6)
            k = d['Kernel Name'][:90]
            tot[k] += v; cnt[k] += 1
print("total ms", sum(tot.values()))
for k, v in tot.most_common(25):
    print(f"{v:9.3f} ms  x{cnt[k]:4d}  {k}")
PY
