timeout 80 python tools/s2d_trace.py > gpurun_out/r02_s2d_trace4.txt 2>&1; grep "===" gpurun_out/r02_s2d_trace4.txt
UPDATE=0 TICKS=3 NOISE_COUNT=250000000 timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_s2d.csv python tools/one_tick.py > gpurun_out/r02_onetick.log 2>&1
timeout 300 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 200 gpurun_out/r02_bench_n1.err
B="--no-e2e --no-cpu-baseline --steps 2 --warmup 3"
DNE_BENCH_BREAKDOWN=1 timeout 120 python bench.py $B --pop 126 > gpurun_out/x_bd126.json 2> gpurun_out/x_bd126.err; grep breakdown gpurun_out/x_bd126.err | tail -2
timeout 120 python bench.py $B --pop 126 > gpurun_out/x_pop126.json 2>/dev/null
DNE_BENCH_T=300 DNE_BENCH_STREAMS=2 timeout 120 python bench.py $B > gpurun_out/x_s2.json 2>/dev/null
DNE_BENCH_T=300 DNE_BENCH_STREAMS=2 DNE_BENCH_PHASED=1 timeout 120 python bench.py $B > gpurun_out/x_s2p.json 2>/dev/null
DNE_BENCH_T=300 timeout 120 python bench.py $B --slots 512 > gpurun_out/x_512.json 2>/dev/null
DNE_BENCH_T=300 timeout 120 python bench.py $B --slots 1024 > gpurun_out/x_1024.json 2>/dev/null
DNE_BENCH_T=300 timeout 120 python bench.py $B > gpurun_out/x_base.json 2>/dev/null
for f in x_pop126 x_base x_s2 x_s2p x_512 x_1024; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"]), d["ms_per_step"], d["roofline"].get("whole_run_frac"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
