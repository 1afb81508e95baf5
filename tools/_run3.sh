timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_gpu_multi.py 2>&1 | tail -15
timeout 200 python tools/vbn_time.py
SLOTS=64 MODES=2 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r02_launches_vbn.csv python tools/vbn_time.py > gpurun_out/vbn_ncu.log 2>&1
