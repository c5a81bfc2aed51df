export XEE_NO_BUILD=1
test -f xlab_ee_fortran_b200/lib/libxee_b200.so || { echo NO_SO; exit 9; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2_weak.json 2> gpurun_out/n2a.err; tail -c 300 gpurun_out/n2a.err
timeout 300 $T bench.py --gpus 2 --steps 3 --warmup 2 --total 4096 > gpurun_out/r02_bench_n2_total4096.json 2> gpurun_out/n2b.err; tail -c 300 gpurun_out/n2b.err
timeout 300 $T bench.py --gpus 2 --steps 3 --warmup 2 --workload series --method line2_chebyshev > gpurun_out/r02_bench_n2_series.json 2> gpurun_out/n2c.err; tail -c 300 gpurun_out/n2c.err
timeout 200 $T bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r02_ref_n2.json 2> gpurun_out/n2d.err
python - <<PY
import json
for f in ("r02_bench_n2_weak","r02_bench_n2_total4096","r02_bench_n2_series","r02_ref_n2"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d.get("n_gpus"), round(d["value"],3), d.get("scaling"), "e2e", d["e2e"] and round(d["e2e"]["value"],3), d.get("config",{}).get("workload","")[:80])
    except Exception as e: print(f, "ERR", e)
PY
