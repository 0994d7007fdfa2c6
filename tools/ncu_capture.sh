#!/bin/bash
# ncu evidence for the bench's headline launch (one GPU, never under torchrun):
#   TAG_launches.csv        per-launch durations of `bench.py --steps 2 --warmup 3` (the share of every kernel in a step)
#   TAG_scan_small.ncu-rep  --set full capture of the full-interval launch of scan_small_kernel<12,0,0>
# usage: tools/ncu_capture.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-companions > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "launch list rc=$?"
# the verify leg launches the kernel on 1 sweep first; launch #2 is a whole 256-sweep interval (2 GiB)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_small_kernel --launch-skip 1 --launch-count 1 \
    -o gpurun_out/${TAG}_scan_small -f python bench.py --steps 2 --warmup 3 --no-cpu --no-companions > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/${TAG}_scan_small.ncu-rep
