# compute-sanitizer on the kernels touched in this session (small cases: the tools are 10-100x slower)
mkdir -p gpurun_out
run() { # name tool pytest-k
  timeout 900 compute-sanitizer --tool $2 python -m pytest tests/test_gpu_parity.py -x -q -k "$3" > gpurun_out/sanitize_$1_$2.log 2>&1
  echo "== $1 $2"; tail -3 gpurun_out/sanitize_$1_$2.log | cut -c1-160
}
run small memcheck "u8_path_all_sizes or survey_known"
run small racecheck "u8_path_all_sizes and (12- or 11- or 8-)"
run small synccheck "u8_path_all_sizes and (12- or 10-)"
run stream memcheck "boxcar_stream_kernel_forced and (8-13 or 10-28 or 12-16) and not 0-"
run stream synccheck "boxcar_stream_kernel_forced and (10-28 or 12-16) and not 0-"
run hb memcheck "fifth_order_chain_every_depth or (fifth_order_streaming and 4-9)"
run hb racecheck "fifth_order_chain_every_depth"
