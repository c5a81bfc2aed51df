export XEE_NO_BUILD=1
test -f xlab_ee_fortran_b200/lib/libxee_b200.so || { echo NO_SO; exit 9; }
for k in 1 2; do
timeout 200 python bench.py --steps 3 --warmup 1 --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('line2 edge-split', round(d['value'],1), round(d['roofline']['avg_launch_us'],1), round(d['roofline']['frac'],3), d['roofline']['sweeps_per_solve'], d['clocks']['sm_mhz'])"
done
timeout 300 python -m pytest tests/test_gpu_twolevel.py tests/test_gpu_line.py -x -q 2>&1 | tail -3
