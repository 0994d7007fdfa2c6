timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -k "async_report or bench_two_ranks or equal_run" 2>&1 | tail -3
python tools/ab_small.py cfg5:4 cfg5:8 cfg5:16 cfg5:32 cfg5:64 cfg5:128 cfg5:256 cfg5n8 cfg5n4 rms1k rms rms16k 2>&1 | grep -v Warn | tee gpurun_out/r02j_sizes.txt
for ns in 1000 2300 4600; do echo "stagger $ns"; RTLSDR_GPU_STAGGER_NS=$ns python tools/ab_small.py cfg5n8 cfg5:256 cfg2 2>&1 | grep -v Warn; done | tee gpurun_out/r02j_stagger.txt
