python tools/ab_small.py rms1k rms rms16k 2>&1 | grep -v Warn
echo "warp-per-read kernel:"; RTLSDR_GPU_RMS_WARP=1 python tools/ab_small.py rms rms16k 2>&1 | grep -v Warn
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -k "rms or fuzz or mixed_submission" 2>&1 | tail -2
