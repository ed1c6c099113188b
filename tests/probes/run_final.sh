#!/bin/bash
# final verification on one B200 (under gpurun): build check is done in the container; here: smoke(), the whole -m gpu suite,
# the driver's default bench command, the reference-on-CUDA datapoint, a clean launch list of the default command
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_final.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_final.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_default.json'));print('default',d['value'],d['e2e'],d['roofline'],d['gpu_launches'])"
timeout 300 python bench.py --impl reference --ref-device cuda --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_cuda.json 2> gpurun_out/r02_bench_reference_cuda.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_reference_cuda.json'));print('ref cuda',d['value'],d['ms_per_step'])"
WRNN_STREAM_DRAWS=0 ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_launches_default.csv \
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_default.log 2>&1
tail -2 gpurun_out/r02_launches_default.log | cut -c1-200
