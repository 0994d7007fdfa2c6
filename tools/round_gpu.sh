# one GPU-box call: parity suite, bench line, ncu launch list of the bench command, full capture of the dominant kernel,
# sanitizer passes on the newest kernel
mkdir -p gpurun_out
TAG=${TAG:-r01o}
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1; grep -h "passed\|failed" gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 200 gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-companions --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
for tool in memcheck synccheck; do
  RTLSDR_GPU_BOXCAR_STREAM=5 timeout 600 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -x -q -k "boxcar_stream_kernel_forced and 5- and (8-13 or 10-28 or 9-28)" > gpurun_out/sanitize_sym_$tool.log 2>&1
  echo "== sym $tool"; tail -3 gpurun_out/sanitize_sym_$tool.log | cut -c1-160
done
