#!/bin/bash
# r02 evidence capture (run under gpurun, 1 GPU): `gpurun --timeout 1500 -- bash profiles/r02_capture.sh`.
# Everything lands in gpurun_out/; the summaries that are judged are copied into profiles/ by hand afterwards.
# Numbers printed by a run under ncu are never bench values.
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"

# 1. bench lines (CUDA events, no profiler)
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r02_bench_default.json 2> $O/r02_bench_default.err
timeout 400 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg3.json 2> $O/r02_bench_cfg3.err
timeout 400 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg4.json 2> $O/r02_bench_cfg4.err
WRNN_STREAM_MIN_FOLDS=65 timeout 400 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg4_stream.json 2> $O/r02_bench_cfg4_stream.err
timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg5_n1.json 2> $O/r02_bench_cfg5_n1.err

# 2. launch lists of the same commands (per-launch durations: the kernel's share of the step)
WRNN_STREAM_DRAWS=0 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02_launches_default.csv \
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r02_launches_default.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02_launches_cfg5.csv \
  timeout 600 python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e > $O/r02_launches_cfg5.log 2>&1

# 3. ncu --set full: the persistent kernel on the whole cfg2 launch, the stream kernel on 40 steps of cfg5
$NCU --set full --import-source on -k regex:wrnn_tc_kernel -c 1 -f -o $O/r02_tc_full \
  timeout 900 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --skip-e2e > $O/r02_tc_full.log 2>&1
$NCU --set full --import-source on -k regex:wrnn_stream_kernel -c 1 -f -o $O/r02_stream_full \
  timeout 900 python bench.py --workload cfg5 --engine stream --seg-steps 40 --steps 1 --warmup 0 --no-cpu-baseline --skip-e2e > $O/r02_stream_full.log 2>&1
$NCU --set full --import-source on -k regex:wrnn_stream_kernel -c 1 -f -o $O/r02_stream_x4_full \
  timeout 900 python bench.py --workload cfg5 --cfg5-folds 512 --engine stream --seg-steps 40 --steps 1 --warmup 0 --no-cpu-baseline --skip-e2e > $O/r02_stream_x4_full.log 2>&1
for f in r02_tc_full r02_stream_full r02_stream_x4_full; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/$f.raw.csv 2>/dev/null
done
ls -la $O | tail -20
