export XEE_NO_BUILD=1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
XEE_TRACE=1 timeout 300 $T bench.py --gpus 8 --steps 2 --warmup 1 --e2e-steps 0 --no-cpu --workload series --method line2_chebyshev > gpurun_out/r02_n8_series_trace.json 2> gpurun_out/n8_trace.err
grep -E "xee trace|subsampled" gpurun_out/n8_trace.err | sort | uniq -c | sort -k5 -n | tail -40
nproc; free -g | head -2
