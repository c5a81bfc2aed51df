export XEE_NO_BUILD=1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T bench.py --gpus 8 --steps 3 --warmup 2 --workload series --method line2_chebyshev > gpurun_out/r02_bench_n8_series.json 2> gpurun_out/n8c.err; tail -c 300 gpurun_out/n8c.err
timeout 300 $T bench.py --gpus 8 --steps 2 --warmup 1 --e2e-steps 0 --workload series --method line_chebyshev > gpurun_out/r02_bench_n8_series_one_level.json 2> gpurun_out/n8e.err; tail -c 300 gpurun_out/n8e.err
python - <<PY
import json
for f in ("r02_bench_n8_series","r02_bench_n8_series_one_level"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d.get("n_gpus"), round(d["value"],1), "ms/step", round(d["ms_per_step"],1), "e2e", d["e2e"] and round(d["e2e"]["value"],1), d["roofline"]["sweeps_per_solve"], d["solves_stopped_on_roundoff_floor"])
    except Exception as e: print(f, "ERR", e)
PY
