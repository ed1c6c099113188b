#!/bin/bash
# quick bench first (keep / revert decision), then the whole -m gpu suite on the same build
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_poll.json 2> gpurun_out/r02_bench_poll.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_poll.json'));print('default',d['value'],d['us_per_sequential_step'],d['e2e']['value'],d['e2e']['ms_per_step'])"
timeout 200 python bench.py --gpus 1 --workload cfg3 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_poll_cfg3.json 2> gpurun_out/r02_bench_poll_cfg3.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_poll_cfg3.json'));print('cfg3',d['value'],d['us_per_sequential_step'],d['e2e']['value'])"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_final2.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_final2.log
