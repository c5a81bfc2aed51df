export XEE_NO_BUILD=1
timeout 300 python -m pytest tests/test_gpu_twolevel.py -x -q 2>&1 | tail -4
for v in "" _v0 ""; do
  export XEE_SO=$PWD/xlab_ee_fortran_b200/lib/libxee_b200$v.so
  timeout 200 python bench.py --steps 3 --warmup 1 --no-cpu --e2e-steps 0 --no-streaming 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('line2 variant[$v]', round(d['value'],1), round(d['roofline']['avg_launch_us'],1), round(d['roofline']['frac'],3), d['roofline']['sweeps_per_solve'], d['clocks']['sm_mhz'])"
done
