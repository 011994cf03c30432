#!/bin/bash
mkdir -p gpurun_out
python tools/train_step_once.py 8 3 > gpurun_out/r2e_train_once.txt 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r2e_launches_train.csv \
    python tools/train_step_once.py 8 1 > gpurun_out/r2e_ncu_train.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2e_launches_train.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
ki, vi = rows[hdr].index('Kernel Name'), rows[hdr].index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    name = r[ki].split('(')[0][:70]
    agg[name][0] += 1; agg[name][1] += float(r[vi].replace(',', '')) / 1e3
tot = sum(v[1] for v in agg.values())
with open('gpurun_out/r2e_train_kernel_shares.txt', 'w') as f:
    f.write(f"total {tot/1e3:.1f} ms over {sum(v[0] for v in agg.values())} launches\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        f.write(f"{v[1]/1e3:9.2f} ms {100*v[1]/tot:5.1f}% {v[0]:6d}  {k}\n")
print(open('gpurun_out/r2e_train_kernel_shares.txt').read())
PY
cat gpurun_out/r2e_train_once.txt
