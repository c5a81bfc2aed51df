export XEE_NO_BUILD=1
test -f xlab_ee_fortran_b200/lib/libxee_b200.so || { echo NO_SO; exit 9; }
timeout 400 python -m pytest tests/test_gpu_twolevel.py -x -q 2>&1 | tail -5
XEE_TRACE=1 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu --method line2_chebyshev > gpurun_out/r02_bench_line2d.json 2> gpurun_out/r02_bench_line2d.err; grep -E "estimate_rho|two-level setup|e2e step" gpurun_out/r02_bench_line2d.err | tail -4
python - <<PY
import json
for k in ("line2d",):
    try:
        d=json.load(open("gpurun_out/r02_bench_%s.json"%k)); print(k, d["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["roofline"]["sweeps_per_solve"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["ms_per_step"], d["clocks"])
    except Exception as e: print(k, "ERR", e)
PY
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --method line2_chebyshev > gpurun_out/r02_plain_line2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"sweep_line|coarse" -s 1500 -c 40 --csv --log-file gpurun_out/r02_launches_line2d.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --method line2_chebyshev > gpurun_out/r02_ncu_line2.log 2>&1
grep -E "sweep_line|coarse" gpurun_out/r02_launches_line2d.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | sort | uniq -c | sort -rn | head -12
