export XEE_NO_BUILD=1
test -f xlab_ee_fortran_b200/lib/libxee_b200.so || { echo NO_SO; exit 9; }
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --method line2_chebyshev > gpurun_out/r02_plain_line2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_line_kernel -s 900 -c 1 -o gpurun_out/r02_prof_line2 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --method line2_chebyshev > gpurun_out/r02_ncu_line2.log 2>&1
tail -3 gpurun_out/r02_ncu_line2.log
