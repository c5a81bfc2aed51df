export XEE_NO_BUILD=1
for v in "" _v1 _v2 _v3 ""; do
  export XEE_SO=$PWD/xlab_ee_fortran_b200/lib/libxee_b200$v.so
  test -f $XEE_SO || { echo NO_SO $v; continue; }
  timeout 200 python bench.py --steps 3 --warmup 1 --no-cpu --e2e-steps 0 --method line2_chebyshev 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant[$v]', round(d['value'],1), round(d['roofline']['avg_launch_us'],1), d['roofline']['sweeps_per_solve'], d['clocks']['sm_mhz'])"
done
