timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_multi.py 2>&1 | tail -15
timeout 200 python tools/vbn_time.py
timeout 120 python tools/tick_time.py
DNE_OPTS=fold_theta=0 timeout 120 python tools/tick_time.py
