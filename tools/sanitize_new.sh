# compute-sanitizer on the kernels added in this session (small cases: the tools are 10-100x slower)
mkdir -p gpurun_out
K='boxcar_stream_kernel_forced and (8-13 or 10-28 or 12-16) and not 0-'
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -x -q -k "$K" > gpurun_out/sanitize_stream_$tool.log 2>&1
  tail -4 gpurun_out/sanitize_stream_$tool.log | cut -c1-200
done
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -x -q -k "fifth_order_chain_every_depth or (fifth_order_streaming and 4-9)" > gpurun_out/sanitize_hb_$tool.log 2>&1
  tail -4 gpurun_out/sanitize_hb_$tool.log | cut -c1-200
done
