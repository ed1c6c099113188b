#!/bin/bash
# 2-GPU checks (under gpurun --gpus 2): the multi-rank GPU tests, then the default bench and cfg4 at N = 2
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_gpu_multi.log 2>&1; tail -3 gpurun_out/r02_gpu_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n2.json'));print('n2',d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e'].get('ms_per_step'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg4 --steps 3 --warmup 3 > gpurun_out/r02_bench_cfg4_n2.json 2> gpurun_out/r02_bench_cfg4_n2.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_cfg4_n2.json'));print('cfg4 n2',d['value'],d['e2e']['value'],d['ms_per_step'])"
