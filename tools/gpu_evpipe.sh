#!/bin/bash
# event pipeline: parity tests, bench line, launch list and one ncu --set full capture of the fused kernel
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_event_pipeline_gpu.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --workload event_pipeline > gpurun_out/bench_evpipe.json 2> gpurun_out/bench_evpipe.err
python -c "
import json; d=json.load(open('gpurun_out/bench_evpipe.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['parity_vs_oracle'], d['cpu_baseline']['value'])"; tail -3 gpurun_out/bench_evpipe.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:event_pipeline_fused -s 5 -c 1 -f -o gpurun_out/evpipe_fused python bench.py --workload event_pipeline --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
