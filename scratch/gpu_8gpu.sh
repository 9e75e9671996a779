#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; tail -3 gpurun_out/bench_8gpu.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_8gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print("N=8 ms_per_step %.3f value %.4g e2e %.4g (%.2f ms; single %.2f) n_gpus %d cpu %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_single_stream"], d["n_gpus"], d["cpu_baseline"]))
PY
