#!/bin/bash
# One GPU-box session: parity tests, smoke, bench (N=1, both arms, all three workloads), ncu launch list + --set full capture
# of exactly one fused step (bench.py's GANTTS_B200_CUDA_PROFILE_STEPS window).
# Usage (from the repo root on the GPU box): bash tools/gpu_session.sh <tag>;  then here: python tools/per_launch.py <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 > gpurun_out/tests_${TAG}.log
tail -2 gpurun_out/tests_${TAG}.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_${TAG}.log 2>&1; tail -2 gpurun_out/smoke_${TAG}.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 300 gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg3.json 2>> gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --workload cfg5 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg5.json 2>> gpurun_out/bench_${TAG}.err
export GANTTS_B200_CUDA_PROFILE_STEPS=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_list_${TAG}.log 2>&1
timeout 1200 ncu --profile-from-start off --set full --clock-control none -o gpurun_out/prof_step_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_full_${TAG}.log 2>&1
unset GANTTS_B200_CUDA_PROFILE_STEPS
# gpurun copies back at most 64 MiB: keep the raw-page CSV of the capture, drop the report itself when it is large
ncu -i gpurun_out/prof_step_${TAG}.ncu-rep --page raw --csv > gpurun_out/step_full_${TAG}.csv 2>/dev/null
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/prof_step_${TAG}.ncu-rep; fi
du -sm gpurun_out
ls -la gpurun_out | grep ${TAG}
