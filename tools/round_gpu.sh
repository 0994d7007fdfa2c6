# one GPU-box call: parity suite, bench line, ncu launch list of the bench command, full captures of the dominant kernels
mkdir -p gpurun_out
TAG=${TAG:-r01m}
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log | head -2
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-companions --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:scan_small_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_scan_small python bench.py --steps 3 --warmup 3 --no-companions --no-cpu > gpurun_out/${TAG}_ncu_small.log 2>&1
ls -la gpurun_out | grep ${TAG}
