bash tools/exp_large.sh r03g 2>&1
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -k "merge_device or read_sharded" 2>&1 | tail -15
