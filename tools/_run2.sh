timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_scale.py tests/test_gpu_parity.py tests/test_gpu_drivers.py -x -q -m gpu 2>&1 | tail -5
timeout 120 python tools/tick_time.py
DNE_OPTS=pdl=0 timeout 120 python tools/tick_time.py
UPDATE=0 TICKS=3 NOISE_COUNT=250000000 timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r02_launches_warm.csv python tools/one_tick.py > gpurun_out/r02_onetick_warm.log 2>&1
