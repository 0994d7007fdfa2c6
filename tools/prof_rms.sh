mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:rms_kernel -s 3 -c 1 -f -o gpurun_out/r01o_rms python tools/scan_bench.py --range 100M:110M:1M --passes 4096 --steps 3 --no-kernel-time > gpurun_out/r01o_ncu_rms.log 2>&1
tail -1 gpurun_out/r01o_ncu_rms.log | cut -c1-100
