# compute-sanitizer on the kernels of this session: vector round C, pipelined tile-major round B (cp.async double
# buffer + per-tile twiddle tables), ticketed byte-sum kernel, merge_sets_kernel.  Small cases: the tools are 10-100x slower.
mkdir -p gpurun_out
run() { # name tool file pytest-k
  timeout 600 compute-sanitizer --tool $2 python -m pytest $3 -x -q -k "$4" > gpurun_out/sanitize_r03_$1_$2.log 2>&1
  echo "== $1 $2"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_r03_$1_$2.log | tail -3 | cut -c1-160
}
run large memcheck tests/test_gpu_parity.py "large_fft_path and (13- or 16- or 17- or 18-)"
run large racecheck tests/test_gpu_parity.py "large_fft_path and (17- or 18-)"
run large synccheck tests/test_gpu_parity.py "large_fft_path and (17- or 18-)"
run merge memcheck tests/test_gpu_round2.py "merge_device"
timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -k "cli_workers" 2>&1 | tail -3
