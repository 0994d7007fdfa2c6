python tools/ab_small.py cfg4 2>&1 | grep -v Warn
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02z_large_launches.csv python tools/ab_small.py cfg4 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r02z_large_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=i;break
h=rows[hdr]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); gs=h.index('Grid Size')
for r in rows[hdr+2:][-4:]:
    print(r[kn][:50], r[gs], r[mv])
PY
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -x -q -k "large or config4 or cfg4" 2>&1 | tail -2
