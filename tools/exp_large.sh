# large-FFT path (config 4): A/B of the pipelined rounds (RTLSDR_GPU_LARGE_PIPE bit 0 = round A, bit 1 = round B),
# per-kernel times of the default under ncu, then the large-path parity tests.  usage: tools/exp_large.sh TAG
TAG=${1:-r03b}
mkdir -p gpurun_out
for m in 0 1; do
  echo "RTLSDR_GPU_LARGE_PIPE=$m"; RTLSDR_GPU_LARGE_PIPE=$m python tools/ab_small.py cfg4 2>&1 | grep -v Warn
done | tee gpurun_out/${TAG}_large_ab.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_large_launches.csv python tools/ab_small.py cfg4 > /dev/null 2>&1
python - <<PY | tee -a gpurun_out/${TAG}_large_ab.txt
import csv
rows=list(csv.reader(open('gpurun_out/${TAG}_large_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=i;break
h=rows[hdr]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); gs=h.index('Grid Size')
for r in rows[hdr+2:][-5:]:
    print(r[kn][:50], r[gs], r[mv])
PY
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -x -q -k "large or config4 or cfg4 or bin_e" 2>&1 | tail -2
