#!/bin/bash
# the whole -m gpu suite on one B200 (under gpurun), log into gpurun_out/
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_final.log 2>&1; tail -4 gpurun_out/r02_gpu_tests_final.log
timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg4.json 2> gpurun_out/r02_bench_cfg4.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_cfg4.json'));print('cfg4',d['value'],d['e2e']['value'],d['config'].get('workload'),d.get('impl_details',{}).get('engine'))"
