export XEE_NO_BUILD=1
test -f xlab_ee_fortran_b200/lib/libxee_b200.so || { echo NO_SO; exit 9; }
timeout 400 python -m pytest tests/test_gpu_twolevel.py -x -q 2>&1 | tail -8
timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu --method line2_chebyshev > gpurun_out/r02_bench_line2b.json 2> gpurun_out/r02_bench_line2b.err; tail -c 400 gpurun_out/r02_bench_line2b.err
python - <<PY
import json
for k in ("line2b",):
    try:
        d=json.load(open("gpurun_out/r02_bench_%s.json"%k)); print(k, d["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["roofline"]["sweeps_per_solve"], d["e2e"], d["clocks"])
    except Exception as e: print(k, "ERR", e)
PY
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --method line2_chebyshev > gpurun_out/r02_plain_line2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep_line|coarse|prolong|finalize" -s 1700 -c 300 --csv --log-file gpurun_out/r02_launches_line2.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --method line2_chebyshev > gpurun_out/r02_ncu_line2.log 2>&1
