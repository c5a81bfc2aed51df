export XEE_NO_BUILD=1
test -f xlab_ee_fortran_b200/lib/libxee_b200.so || { echo NO_SO; exit 9; }
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; tail -c 300 gpurun_out/r02_bench_n1_final.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n1_final.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), "us/sweep", round(d["roofline"]["avg_launch_us"],1), d["roofline"]["sweeps_per_solve"], d["clocks"])
s=d["roofline_streaming_kernel"]; print("streaming", round(s["frac"],3), round(s["avg_launch_us"],1), round(s["solves_per_s"],1), s["sweeps_per_solve"])
PY
