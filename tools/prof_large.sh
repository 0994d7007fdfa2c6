# ncu --set full of the large-path kernels of config 4 (one step after warm-up).  Only text travels back:
# the raw-page summary and the per-instruction source page of rounds A and B (gpurun_out/ is capped at 64 MiB per call).
TAG=${1:-r03j}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"large_round|dc_sums" --launch-skip 12 --launch-count 4 -f -o /tmp/${TAG}_large python tools/ab_small.py cfg4 > gpurun_out/${TAG}_ncu.log 2>&1
echo "rc=$?"
python tools/ncu_summary.py /tmp/${TAG}_large.ncu-rep > gpurun_out/${TAG}_large_ncu_summary.txt
ncu -i /tmp/${TAG}_large.ncu-rep --page source --csv --kernel-name regex:large_round_a > gpurun_out/${TAG}_round_a_source.csv 2>/dev/null
ncu -i /tmp/${TAG}_large.ncu-rep --page source --csv --kernel-name regex:large_round_b > gpurun_out/${TAG}_round_b_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*
