mkdir -p gpurun_out
CFG=${CFG:-10:28}
P=$(python -c "
be,ds=map(int,'$CFG'.split(':')); print(max(64,(400<<20)//(2*(1<<be)*ds)))")
RTLSDR_GPU_BOXCAR_STREAM=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:scan_boxcar_stream -s 3 -c 1 -f -o gpurun_out/s2_stream_prof python tools/scan_bench.py --boxcar $CFG --passes $P --steps 3 --no-kernel-time > gpurun_out/s2_ncu.log 2>&1
tail -2 gpurun_out/s2_ncu.log
