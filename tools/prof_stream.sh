mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "boxcar_stream_kernel_forced or fused_boxcar_path" > gpurun_out/s2_stream_pytest.log 2>&1
tail -3 gpurun_out/s2_stream_pytest.log
RTLSDR_GPU_BOXCAR_STREAM=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:scan_boxcar_stream -s 3 -c 1 -f -o gpurun_out/s2_stream_1028 python tools/scan_bench.py --boxcar 10:28 --passes 7314 --steps 3 --no-kernel-time > gpurun_out/s2_ncu.log 2>&1
tail -3 gpurun_out/s2_ncu.log
ls -la gpurun_out/
