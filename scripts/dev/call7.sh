export XEE_NO_BUILD=1
test -f xlab_ee_fortran_b200/lib/libxee_b200.so || { echo NO_SO; exit 9; }
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --method line2_chebyshev > gpurun_out/r02_plain_line2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep_line|coarse|prolong|finalize" -s 800 -c 400 --csv --log-file gpurun_out/r02_launches_line2.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --method line2_chebyshev > gpurun_out/r02_ncu_line2.log 2>&1
XEE_TRACE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --method line2_chebyshev 2>&1 | grep -E "xee trace|two-level|e2e step|spectral" | tail -40
