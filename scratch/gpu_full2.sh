#!/bin/bash
# full GPU suite + default bench line + 2-rank torchrun bench (needs --gpus 2)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print("N=1 ms_per_step %.3f value %.4g e2e %.4g (%.2f ms) launches %d clocks %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"]))
print(d["roofline"]); print(d["cpu_baseline"])
PY
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -2 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print("N=2 ms_per_step %.3f value %.4g e2e %.4g n_gpus %d"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["n_gpus"]))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | tail -1 | cut -c1-300
fi
