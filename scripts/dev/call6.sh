export XEE_NO_BUILD=1
test -f xlab_ee_fortran_b200/lib/libxee_b200.so || { echo NO_SO; exit 9; }
timeout 400 python -m pytest tests/test_gpu_twolevel.py -x -q -s 2>&1 | grep -v "^$" | tail -30
timeout 500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_twolevel.py 2>&1 | tail -30
timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu --method line2_chebyshev > gpurun_out/r02_bench_line2.json 2> gpurun_out/r02_bench_line2.err; tail -c 400 gpurun_out/r02_bench_line2.err
timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r02_bench_line1.json 2> gpurun_out/r02_bench_line1.err
python - <<PY
import json
for k in ("line2","line1"):
    try:
        d=json.load(open("gpurun_out/r02_bench_%s.json"%k)); print(k, d["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["roofline"]["sweeps_per_solve"], d["e2e"]["value"], d["clocks"])
    except Exception as e: print(k, "ERR", e)
PY
