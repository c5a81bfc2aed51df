export XEE_NO_BUILD=1
test -f xlab_ee_fortran_b200/lib/libxee_b200.so || { echo NO_SO; exit 9; }
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_default.json 2> gpurun_out/r02_bench_n1_default.err; tail -c 300 gpurun_out/r02_bench_n1_default.err
timeout 400 python bench.py --steps 10 --warmup 3 --method line_chebyshev --no-cpu > gpurun_out/r02_bench_n1_line_chebyshev.json 2> /dev/null
# ncu full of the two shipped sweep kernels (steady-state launches), then the launch lists
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/plain_a.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_line_kernel -s 1200 -c 2 -o gpurun_out/r02_prof_line2_final python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/ncu_a.log 2>&1
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --method line_chebyshev > gpurun_out/plain_b.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_line_kernel -s 1500 -c 2 -o gpurun_out/r02_prof_line1_final python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --method line_chebyshev > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 2500 --csv --log-file gpurun_out/r02_launches_line2_step.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 2600 --csv --log-file gpurun_out/r02_launches_line1_step.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --method line_chebyshev > gpurun_out/ncu_d.log 2>&1
echo done
