#!/bin/bash
# experiments (under gpurun): cluster form of the stream engine, timing per fold count + cycle counters of CTA 0
timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -x -q -k "${SWEEP_K:-stream}" > gpurun_out/r02_t12.log 2>&1; tail -3 gpurun_out/r02_t12.log
for folds in ${SWEEP_FOLDS:-16 512 1024}; do for cl in 4; do
  WRNN_STREAM_CL=$cl WRNN_STREAM_PROF=1 timeout 120 python bench.py --workload cfg5 --cfg5-folds $folds --engine stream --seg-steps 300 --skip-e2e --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/sw_${folds}_$cl.json 2> gpurun_out/sw_${folds}_$cl.err
  python -c "import json;d=json.load(open('gpurun_out/sw_${folds}_$cl.json'));print($folds,$cl,round(d['value']/1e6,2),round(d['us_per_sequential_step'],1),d.get('impl_details',{}).get('engine'))" 2>&1 | tail -1
  grep "wrnn_stream prof" gpurun_out/sw_${folds}_$cl.err | tail -1
done; done
