# quick check of the shipped library after a series of reverted experiments: smoke + a cross-section of the GPU suite
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_baseline_configs.py -x -q -k "large or fifth_order or merge_device or survey_known or compiled_reference or rms" 2>&1 | tail -2
python tools/ab_small.py cfg5 cfg4 f9 2>&1 | grep -v Warn
