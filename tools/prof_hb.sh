mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"halfband_stream|halfband_chain" -s 8 -c 2 -f -o gpurun_out/s2_hb python tools/scan_bench.py --range 100M:100.1M:100 -F 9 -w blackman --passes 4096 --steps 3 --no-kernel-time > gpurun_out/s2_ncu_hb.log 2>&1
tail -2 gpurun_out/s2_ncu_hb.log
