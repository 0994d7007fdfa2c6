# ncu launch list of the bench command (per-launch durations; the transform's share of a step)
TAG=${1:-r03n}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-companions > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "launch list rc=$?"
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/${TAG}_launches.csv")))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr = i
        break
h = rows[hdr]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[hdr + 2:]:
    if len(r) > mv:
        tot[r[kn][:60]] += float(r[mv].replace(",", "")); cnt[r[kn][:60]] += 1
for k, v in tot.most_common(8):
    print(f"{k:60s} n={cnt[k]:4d} total={v/1e3:10.1f} us")
PY
