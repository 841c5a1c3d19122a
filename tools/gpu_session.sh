#!/bin/bash
# One GPU-box session: parity tests, smoke, bench (N=1, both arms), ncu launch list + full capture of the GEMMs.
# Usage (from the repo root on the GPU box): bash tools/gpu_session.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_${TAG}.csv &
SMI=$!
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
kill $SMI
tail -c 400 gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_${TAG}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 690 -c 140 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 46 -c 31 -o gpurun_out/prof_gemm_${TAG} python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
