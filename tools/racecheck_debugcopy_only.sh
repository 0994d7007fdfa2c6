TAG=${1:-r02q}
DBG=build/racecheck/librtlsdr_gpu_scan_racecopy.so
K="boxcar_stream_kernel_forced and (10-28 or 12-16) and (5-0 or 1-0 or 3-1)"
RTLSDR_GPU_SCAN_LIB=$PWD/$DBG timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -k "$K" > gpurun_out/${TAG}_racecheck_debugcopy.txt 2>&1
echo "== debug-copy build"; tail -4 gpurun_out/${TAG}_racecheck_debugcopy.txt | cut -c1-200
