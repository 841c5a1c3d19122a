mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -25 > gpurun_out/t14.log
tail -3 gpurun_out/t14.log
python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b14.json 2> gpurun_out/b14.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/b14.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['gpu_launches_per_step'], d['roofline']['gemm_family'], d['roofline']['traffic'])
PY
export GANTTS_B200_CUDA_PROFILE_STEPS=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_list_r02c.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r02c.csv gemm | tail -30
