export XEE_NO_BUILD=1
test -f xlab_ee_fortran_b200/lib/libxee_b200.so || { echo NO_SO; exit 9; }
timeout 400 python -m pytest tests/test_gpu_twolevel.py tests/test_gpu_series.py -x -q 2>&1 | tail -4
for m in line2_chebyshev line_chebyshev; do
XEE_TRACE=1 timeout 300 python bench.py --workload series --steps 3 --warmup 2 --no-cpu --method $m > gpurun_out/r02_series_$m.json 2> gpurun_out/r02_series_$m.err
grep -E "estimate_rho|two-level setup|subsampled" gpurun_out/r02_series_$m.err | tail -4
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_series_$m.json")); r=d["roofline"]; print("$m", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "us/sweep", round(r["avg_launch_us"],1), "frac", round(r["frac"],3), r["sweeps_per_solve"], "probe share", round(r.get("spectral_probe_share_of_step",0),3), d["clocks"]["sm_mhz"])
except Exception as e: print("$m ERR", e)
PY
done
XEE_TRACE=1 timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 2 2> gpurun_out/e.err | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('map line2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['roofline']['avg_launch_us'],1))"
grep -E "two-level setup|estimate_rho" gpurun_out/e.err | tail -2
