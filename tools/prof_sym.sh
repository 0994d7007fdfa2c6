mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:scan_boxcar_sym -s 3 -c 1 -f -o gpurun_out/r01o_sym_n512_ds28 python tools/scan_bench.py --boxcar 9:28 --passes 14628 --steps 3 --no-kernel-time > gpurun_out/r01o_ncu_sym.log 2>&1
tail -2 gpurun_out/r01o_ncu_sym.log | cut -c1-200
