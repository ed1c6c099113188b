#!/bin/bash
# 8-GPU bench lines (under gpurun --gpus 8): default workload (weak scaling) and cfg5 (4096 folds sharded 8 ways), both with e2e
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n8.json'));print('n8',d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e'].get('ms_per_step'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload cfg5 --steps 3 --warmup 2 > gpurun_out/r02_bench_cfg5_n8.json 2> gpurun_out/r02_bench_cfg5_n8.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_cfg5_n8.json'));print('cfg5 n8',d['value'],d['e2e']['value'],d['ms_per_step'],d.get('impl_details',{}).get('engine'))"
