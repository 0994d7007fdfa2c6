# ncu --set full of the large-path kernels of config 4 (one step after warm-up): pipelined and one-tile-per-CTA rounds.
# Only the text summaries and ONE report travel back (gpurun_out/ is capped at 64 MiB per call).
TAG=${1:-r03c}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:large_round --launch-skip 9 --launch-count 3 -f -o gpurun_out/${TAG}_large_pipe python tools/ab_small.py cfg4 > gpurun_out/${TAG}_ncu_pipe.log 2>&1
echo "pipe rc=$?"
python tools/ncu_summary.py gpurun_out/${TAG}_large_pipe.ncu-rep > gpurun_out/${TAG}_large_pipe_ncu_summary.txt
RTLSDR_GPU_LARGE_PIPE=0 timeout 300 ncu --set full --clock-control none -k regex:large_round --launch-skip 9 --launch-count 2 -f -o /tmp/${TAG}_large_old python tools/ab_small.py cfg4 > gpurun_out/${TAG}_ncu_old.log 2>&1
echo "old rc=$?"
python tools/ncu_summary.py /tmp/${TAG}_large_old.ncu-rep > gpurun_out/${TAG}_large_old_ncu_summary.txt
ls -la gpurun_out/${TAG}_large_*
