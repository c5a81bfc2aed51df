export XEE_NO_BUILD=1
test -f xlab_ee_fortran_b200/lib/libxee_b200.so || { echo NO_SO; exit 9; }
timeout 300 python -m pytest tests/test_gpu_twolevel.py -x -q 2>&1 | tail -3
for k in 1 2; do
timeout 200 python bench.py --steps 3 --warmup 1 --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('line2 splitK', round(d['value'],1), round(d['roofline']['avg_launch_us'],1), round(d['roofline']['frac'],3), d['roofline']['sweeps_per_solve'], d['clocks']['sm_mhz'])"
done
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/r02_plain_line2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep_line|coarse|finalize" -s 1500 -c 200 --csv --log-file gpurun_out/r02_launches_line2e.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/r02_ncu_line2.log 2>&1
