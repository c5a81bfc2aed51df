export XEE_NO_BUILD=1
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_twolevel.py -x -q -k "(jacobi_sweeps and (shape5 or shape3 or shape6)) or one_operator_per_solve" 2>&1 | tail -15
echo "memcheck rc=$?"
