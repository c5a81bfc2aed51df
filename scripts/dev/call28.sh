export XEE_NO_BUILD=1
test -f xlab_ee_fortran_b200/lib/libxee_b200.so || { echo NO_SO; exit 9; }
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/plain_a.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_line_kernel -s 600 -c 2 -o gpurun_out/r02_prof_line2_final python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/ncu_a.log 2>&1
tail -2 gpurun_out/ncu_a.log
timeout 600 ncu --set full --clock-control none -k regex:"coarse_gemm|coarse_gather|coarse_finish" -s 300 -c 3 -o gpurun_out/r02_prof_coarse_final python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/ncu_e.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 3000 --csv --log-file gpurun_out/r02_launches_line2_step.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/ncu_c.log 2>&1
timeout 300 python scripts/run_configs.py --method line2_chebyshev 2 3 5 2>&1 | tail -5
cp gpurun_out/configs.json gpurun_out/r02_configs_2_3_5_line2_chebyshev.json
echo done
