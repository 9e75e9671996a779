#!/bin/bash
# Round-end check on one B200: the full -m gpu suite, smoke(), the default bench line and the launch list of the bench command.
#   gpurun --timeout 1500 -- 'bash tools/final_gpu_check.sh <tag>'
tag=${1:-final}
timeout 1300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 1500 gpurun_out/${tag}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launch_list.csv python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/${tag}_ncu_bench.log 2>&1; tail -1 gpurun_out/${tag}_ncu_bench.log | cut -c1-200
