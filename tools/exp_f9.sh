python tools/ab_small.py f9 2>&1 | grep -v Warn
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -k "fifth_order or decimating or narrow-F9 or survey_known or fuzz" 2>&1 | tail -2
